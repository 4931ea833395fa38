"""Scan synthetic head operating points (gain, bias) on the CPU: fp32 oracle vs its fp16-storage emulation (oracle.nets.forward_fp16_emulated).
Prints the probability / box error the device path will show, the candidates per frame at box_score 0.4 and how many sit within 1e-3 / 2e-3
of the threshold.  Test infrastructure (uses oracle/):  python tools/synth_operating_point.py yolov8 l 45,-9 40,-8.2
"""
import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
torch.set_num_threads(8)
import adas_b200
from adas_b200 import plan
from oracle import nets, post
import synth
kind, variant = sys.argv[1], sys.argv[2]
builder = plan.build_yolov8 if kind=="yolov8" else plan.build_yolov5
x = torch.from_numpy(np.concatenate([post.yolo_prepare_input(synth.frame(s), 640, 640)[0] for s in (0,1,2,3,4,5,6,7)]))
for a in sys.argv[3:]:
    g,b = map(float,a.split(','))
    if kind=="yolov8":
        plan.SYNTH_PROFILES["yolov8"]={"gains":[(r"model\.22\.cv3\.\d\.2\.weight", g), (r"model\.22\.cv2\.\d\.2\.weight", 25.0)],"fill":[(r"model\.22\.cv3\.\d\.2\.bias", b)]}
    else:
        plan.SYNTH_PROFILES["yolov5"]={"gains":[(r"model\.24\.m\.\d\.weight", g)],"fill":[(r"model\.24\.m\.\d\.bias", b)]}
    W = plan.synth_weights(kind, 0); builder(W, variant)
    md = nets.build(kind, W.state_dict, scale=variant)
    with torch.no_grad():
        ref = md(x).numpy(); emu = nets.forward_fp16_emulated(md, x).numpy()
    if kind=="yolov8":
        e=np.abs(ref[:,4:]-emu[:,4:]); mx=ref[:,4:].max(1); mg=emu[:,4:].max(1); eb=np.abs(ref[:,:4]-emu[:,:4]).max()
    else:
        e=np.abs(ref[...,4:]-emu[...,4:]); mx=(ref[...,5:]*ref[...,4:5]).max(2); mg=(emu[...,5:]*emu[...,4:5]).max(2); eb=np.abs(ref[...,:4]-emu[...,:4]).max()
    cand=mx>0.4
    print(f"g={g} b={b}: max prob err {e.max():.2e} box err {eb:.3f} | cands {cand.sum(1).tolist()} within1e-3 {(np.abs(mx-0.4)<1e-3).sum(1).tolist()} within2e-3 {int((np.abs(mx-0.4)<2e-3).sum())} flips {int((cand!=(mg>0.4)).sum())}", flush=True)
    if kind=="yolov8": print("   per-frame max prob err", [f"{v:.2e}" for v in e.max(axis=(1,2))])
    else: print("   per-frame max prob err", [f"{v:.2e}" for v in e.reshape(e.shape[0],-1).max(1)])
