#!/usr/bin/env python
"""Per-CTA timeline of one fused launch on the bench workload (profiling aid; uses B200AugFusedArgs.trace_out).
Prints phase durations by sample class and the SM occupancy picture.  usage: python scripts/trace_ctas.py [out.npz]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200")]
import numpy as np
import torch

import bench
from trackertraincode_b200 import _native as N
from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
from trackertraincode_b200.datatransformation import _engine as E

dev = torch.device("cuda", 0)
host = bench.make_host_batch(0)
gp, pp = bench.draw_params(100, bench.BATCH, 0)
cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
b = Batch(Metadata((bench.SRC, bench.SRC), bench.BATCH, "bench", None, cats), {k: torch.from_numpy(v).to(dev) for k, v in host.items()})
photo = E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma),
                      torch.from_numpy(pp.contrast), torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply),
                      pp.noise_std, pp.seed, pp.sample_offset, pp.clip)
an = torch.from_numpy(gp.angles)
geo = E.GeoParams(torch.from_numpy(gp.scales), an, torch.from_numpy(gp.translations), E.host_cos_sin(an))
flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_PHOTOMETRIC | N.F_WHITEN
call = E.prepare_fused(b, flags=flags, out_size=bench.OUT, geo=geo, do_flip=torch.from_numpy(gp.do_flip.astype(np.uint8)),
                       rot_dir=torch.from_numpy(gp.rot_dir), photo=photo, want_status=True, want_trace=True, want_view_roi=True,
                       cluster_size=int(os.environ.get("B200AUG_CLUSTER", "0")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SPLIT = os.environ.get("B200AUG_TRACE_SPLIT")  # plan phase, device idle, then the main phase alone (no overlap with plan_kernel)
for _ in range(5):
    flush.zero_()
    if SPLIT:
        call.launch_plan()
        torch.cuda.synchronize()
        call.launch_main()
    else:
        call.launch()
torch.cuda.synchronize()
t = call.result.trace.cpu().numpy().astype(np.int64)
CL = int(os.environ.get("B200AUG_CLUSTER", "0")) or 2
live = t[:, 0] > 0
is_worker = live & (t[:, 1] == 0)  # canvas workers (the first CTAs of the grid) never mark "plan loaded"
tw = t[is_worker]
t = t[live & ~is_worker][: bench.BATCH * CL]
t0 = min(t[:, 0].min(), tw[:, 0].min()) if len(tw) else t[:, 0].min()
if len(tw):
    print(f"canvas workers: {len(tw)} CTAs on {len(np.unique(tw[:, 5]))} SMs; start {(tw[:, 0].min() - t0) / 1e3:.1f}..{(tw[:, 0].max() - t0) / 1e3:.1f} us, "
          f"end {(tw[:, 4].min() - t0) / 1e3:.1f}..{(tw[:, 4].max() - t0) / 1e3:.1f}; items per CTA {tw[:, 2].min()}..{tw[:, 2].max()} (total {tw[:, 2].sum()})")
if len(tw):
    nt = max(int(tw[:, 12].sum()), 1)
    print("canvas workers, cycles per tile (thread 0): issue %.0f wait %.0f gather %.0f describe+barrier %.0f; tiles %d; total CTA-us %.0f" % (
        tw[:, 8].sum() / nt, tw[:, 9].sum() / nt, tw[:, 10].sum() / nt, tw[:, 11].sum() / nt, nt, (tw[:, 4] - tw[:, 0]).sum() / 1e3))
if len(tw):
    ni = max(int(tw[:, 7].sum()), 1)
    print("canvas workers, cycles per item (thread 0): fetch %.0f setup %.0f tail %.0f; items %d" % (
        tw[:, 13].sum() / ni, tw[:, 14].sum() / ni, tw[:, 15].sum() / ni, ni))
start, plan, tabs, res, end, smid = (t[:, i] - (t0 if i < 5 else 0) for i in range(6))
print(f"kernel span {(end.max()) / 1e3:.1f} us; CTA duration us: mean {np.mean(end - start) / 1e3:.1f} median {np.median(end - start) / 1e3:.1f} "
      f"p95 {np.percentile(end - start, 95) / 1e3:.1f} max {(end - start).max() / 1e3:.1f}")
_ord = E.launch_order(bench.BATCH, geo, photo)
samp = _ord.numpy().astype(np.int64) if _ord is not None else np.arange(bench.BATCH)  # cluster i of the grid -> sample
rot = np.repeat((gp.angles != 0)[samp], CL)
blur = np.repeat((pp.apply[:, 5] & (5 in list(pp.order)))[samp].astype(bool), CL)
eq = np.repeat((pp.apply[:, 0] & (0 in list(pp.order)))[samp].astype(bool), CL)
noise = np.repeat(pp.noise_apply.any(1)[samp], CL)
warp_stage = t[:, 6] - t0 - tabs
m6, m7, m8, m9, m10 = (t[:, i] - t0 for i in (6, 7, 8, 9, 10))
ph = {"plan": plan - start, "tab": m8 - plan, "cv_wait": m7 - m8, "csync": tabs - m7, "warp": m6 - tabs, "area+xchg": res - m6, "labels": m9 - res,
      "lut": np.where(m10 > 0, m10 - m9, 0), "output": np.where(m10 > 0, end - m10, end - m9), "total": end - start}
for name, m in (("all", np.ones_like(rot)), ("unrotated", ~rot), ("rotated", rot), ("blur", blur), ("equalize", eq), ("noise", noise),
                ("plain(no rot/photo)", ~rot & ~blur & ~eq & ~noise)):
    if m.sum():
        print(f"{name:22s} n={int(m.sum()):4d} " + " ".join(f"{k}={np.mean(v[m]) / 1e3:7.1f}" for k, v in ph.items()))
early = start < 8e3
print(f"first wave: {int(early.sum())} CTAs on {len(np.unique(smid[early]))} SMs, end times us: " + str(np.round(np.sort(end[early])[::8] / 1e3, 1).tolist()))
print("starts (us) sorted, every 32nd:", np.round(np.sort(start)[::32] / 1e3, 1).tolist())
print("start-time histogram (us):", np.histogram(start / 1e3, bins=10)[0].tolist(), "last start", start.max() / 1e3)
busy = np.zeros(int(smid.max()) + 1)
for s, a, e in zip(smid, start, end):
    busy[s] = max(busy[s], e)
sys.stdout.flush()
print(f"per-SM last-CTA end (us): min {busy.min() / 1e3:.1f} median {np.median(busy) / 1e3:.1f} max {busy.max() / 1e3:.1f}; SMs used {len(np.unique(smid))}")
vr = np.repeat(call.result.view_roi.cpu().numpy()[samp], CL, axis=0)
dur = end - start
for i in np.argsort(-dur)[:8]:
    print(f"slow cta {i} (sample {samp[i // CL]}): total {dur[i] / 1e3:.1f} us area {(res - (t[:, 6] - t0))[i] / 1e3:.1f} view_roi {vr[i].tolist()} size {(vr[i, 2] - vr[i, 0], vr[i, 3] - vr[i, 1])} "
          f"rot {bool(rot[i])} blur {bool(blur[i])} eq {bool(eq[i])} noise {bool(noise[i])} sm {smid[i]} start {start[i] / 1e3:.1f}")
if len(sys.argv) > 1:
    np.savez(sys.argv[1], trace=t, rot=rot, blur=blur, eq=eq, noise=noise)
