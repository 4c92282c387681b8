"""Diagnostic: what do envs look like when the fp32 and float64 kernels disagree on a flag?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from grasp_lab_salp_b200 import PRECISION_F64, PRECISION_MIXED, SalpBatch, default_params
from grasp_lab_salp_b200.params import FIELDS
n, T = 4096, 1500
dev = torch.device("cuda", 0)
mixed = SalpBatch(n, default_params(precision=PRECISION_MIXED), seed=0)
f64 = SalpBatch(n, default_params(precision=PRECISION_F64), seed=0)
mixed.reset_device(); f64.reset_device()
sync_cols = [c for c in FIELDS if not (c.startswith("obstacle") and int(c[8]) >= 2)]
sm = {c: mixed.state_tensor(c) for c in sync_cols}; sf = {c: f64.state_tensor(c) for c in sync_cols}
g = torch.Generator(device=dev); g.manual_seed(1234)
rows = []
prev = {}
for t in range(T):
    a = torch.rand((n, 3), generator=g, device=dev); a[:, 2] = a[:, 2] * 2 - 1
    pre = {c: sf[c].clone() for c in ("cycle", "euler_x", "euler_y", "posw_x", "posw_y", "target_x", "target_y")}
    prem = {c: sm[c].clone() for c in ("euler_x", "euler_y", "posw_x", "posw_y")}
    om, rm, tem, trm = mixed.step_device(a, auto_reset=True, extras=True)
    of, rf, tef, trf = f64.step_device(a, auto_reset=True, extras=True)
    bad = (tem != tef) | (trm != trf)
    if bool(bad.any()):
        idx = torch.nonzero(bad).flatten()
        for i in idx.tolist()[:50]:
            rows.append(dict(t=t, cycle=int(pre["cycle"][i]), roll=float(pre["euler_x"][i]), pitch=float(pre["euler_y"][i]),
                             dpos=float(torch.hypot(prem["posw_x"][i]-pre["posw_x"][i], prem["posw_y"][i]-pre["posw_y"][i])),
                             a0=float(a[i,0]), K=int(f64.dev["substeps"][i]),
                             m=(int(tem[i]), int(trm[i])), f=(int(tef[i]), int(trf[i])),
                             blow_m=float(mixed.dev["metrics"][i,19]), blow_f=float(f64.dev["metrics"][i,19]),
                             rm=float(rm[i]), rf=float(rf[i]),
                             tobs_m=[round(float(x),4) for x in mixed.dev["terminal_obs"][i,:2]], tobs_f=[round(float(x),4) for x in f64.dev["terminal_obs"][i,:2]]))
        for c in sync_cols: sm[c][bad] = sf[c][bad]
        mixed.dev["obs"][bad] = f64.dev["obs"][bad]
print(len(rows))
import collections
print("cycle hist:", np.histogram([r["cycle"] for r in rows], bins=[0,5,20,50,100,200,300,400,500])[0])
print("|roll|>0.1:", sum(abs(r["roll"])>0.1 or abs(r["pitch"])>0.1 for r in rows), "dpos>1e-3:", sum(r["dpos"]>1e-3 for r in rows))
print("blow_m:", sum(r["blow_m"]==1 for r in rows), "blow_f:", sum(r["blow_f"]==1 for r in rows))
print("a0 in [0.085,0.096]:", sum(0.085<r["a0"]<0.096 for r in rows))
for r in rows[:25]: print(r)
