#!/usr/bin/env python
"""Resident CTAs per SM over time from a trace saved by scripts/trace_ctas.py (usage: trace_occupancy.py trace.npz ...)."""
import sys

import numpy as np

for path in sys.argv[1:]:
    t = np.load(path)["trace"]
    live = t[:, 0] > 0
    st, en, sm = t[live, 0], t[live, 4], t[live, 5]
    t0 = st.min()
    st, en = (st - t0) / 1e3, (en - t0) / 1e3
    times = np.arange(0.5, en.max(), 1.0)
    conc = np.zeros(len(times))
    peak = []
    for s in np.unique(sm):
        m = sm == s
        c = ((st[m][None, :] <= times[:, None]) & (en[m][None, :] > times[:, None])).sum(1)
        peak.append(c.max())
        conc += c
    n_sm = len(peak)
    print(path, f"span {en.max():.1f} us; CTAs {live.sum()}; CTA-us {float((en - st).sum()):.0f} -> {float((en - st).sum()) / n_sm / en.max():.2f} resident CTAs/SM on average")
    print("  resident CTAs/SM at t = 0.5, 10.5, ... us:", np.round(conc[::10] / n_sm, 2).tolist())
    print("  SMs by peak residency:", dict(enumerate(np.bincount(peak).tolist())))
