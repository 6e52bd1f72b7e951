import sys, time, ctypes as C
sys.path[:0] = ['.', 'beamforming-lk_b200', 'tests']
import numpy as np, torch
import bflk, cases
from bflk import synth
c = cases.CFG4
m = bflk.MISOWorker(cases.origins(c["nx"], c["ny"]))
th, ph = cases.cfg4_targets()
win = synth.make_stream(synth.tile_geometry(cases.origins(c["nx"], c["ny"])), 1024)
dev = torch.device("cuda:0")
wd = torch.from_numpy(win).to(dev)
for T in (16, 104):
    rng = np.random.default_rng(1)
    t = np.ascontiguousarray(rng.random(T) * 1.2); p = np.ascontiguousarray(rng.random(T) * 6.0)
    audio = torch.zeros((T, 256), device=dev); power = torch.zeros(T, device=dev)
    st = torch.cuda.Stream()
    L = m._L
    def call():
        rc = L.bflk_miso_dev(m._h, t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), T, C.c_void_p(wd.data_ptr()), C.c_void_p(audio.data_ptr()), C.c_void_p(power.data_ptr()), C.c_void_p(st.cuda_stream))
        assert rc == 0
    for _ in range(10): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(n): call()
        e1.record(st)
    torch.cuda.synchronize()
    print(f"T={T}: device time per call {e0.elapsed_time(e1)/n*1e3:.1f} us, host wall per call {(time.perf_counter()-t0)/n*1e6:.1f} us")
