Axes3D = type('Axes3D', (), {})
