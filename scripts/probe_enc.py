"""Scratch probe: encoder latency/throughput (L=64) with CUDA events."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
w = os.path.join(tempfile.gettempdir(), "w.sgw")
sg.export_bert_weights(sg.random_minilm_state_dict(1), w, 6, 384, 12, 1536, 30522, 512)
enc = sg.GpuMiniLM(w)
rng = np.random.default_rng(0)
L = int(os.environ.get("L", 64))
ids = torch.from_numpy(rng.integers(1000, 30522, (256, L)).astype(np.int64)).cuda()
mask = torch.ones((256, L), dtype=torch.int64, device="cuda")
out = torch.empty((256, 384), device="cuda")
cs = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for b, reps in ((1, 200), (16, 100), (256, 50)):
    for _ in range(5):
        enc.encode_ids_device(ids.data_ptr(), mask.data_ptr(), b, L, out.data_ptr(), cs)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        enc.encode_ids_device(ids.data_ptr(), mask.data_ptr(), b, L, out.data_ptr(), cs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = (6 * 2 * (4 * 384 * 384 + 2 * 384 * 1536) + 6 * 4 * L * 384) * b * L
    print(f"b={b} L={L}: {ms*1e3:.1f} us  {b/ms*1e3:.0f} q/s  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
