import sys, time, numpy as np
sys.path.insert(0, '.')
import optical_flow_b200 as ofb, synth_frames
from optical_flow_b200.engine import pinned_empty
W,H,P=1920,1080,300
eng=ofb.Farneback(0)
fr=synth_frames.shot(W,H,P+1,seed=5)
pf=pinned_empty(fr.shape,np.uint8); pf[:]=fr
out=pinned_empty((P,H,W,3),np.uint8)
for label,kw in [("bgr",dict(want_bgr=True,out_bgr=out)),("magsum",dict(want_bgr=False,want_magsum=True)),("bgr",dict(want_bgr=True,out_bgr=out))]:
    for i in range(3): r=eng.shot(pf,**kw,**ofb.REFERENCE_PARAMS)
    ms=[eng.shot(pf,**kw,**ofb.REFERENCE_PARAMS)["device_ms"] for i in range(5)]
    print(label, "ms", round(min(ms),2), "pairs/s", round(P/min(ms)*1e3))
# raw copy bandwidth
import ctypes as C
L=eng._L
d=eng.device_alloc(out.nbytes)
for i in range(2): eng.d2h(out,d)
t=time.perf_counter(); eng.d2h(out,d); dt=time.perf_counter()-t
print("D2H pinned GB/s", out.nbytes/dt/1e9)
t=time.perf_counter(); eng.h2d(d,out); dt=time.perf_counter()-t
print("H2D pinned GB/s", out.nbytes/dt/1e9)
