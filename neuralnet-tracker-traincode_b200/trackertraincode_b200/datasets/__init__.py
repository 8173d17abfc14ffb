from .batch import Batch, Metadata, FieldCategory, imagelike_categories  # noqa: F401
