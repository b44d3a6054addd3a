"""Target for the ncu launch list of one encoder forward (b=1 and b=256, L=64)."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shodh_memory_b200 as sg
w = os.path.join(tempfile.gettempdir(), "w.sgw")
sg.export_bert_weights(sg.random_minilm_state_dict(1), w, 6, 384, 12, 1536, 30522, 512)
enc = sg.GpuMiniLM(w)
rng = np.random.default_rng(0)
ids = torch.from_numpy(rng.integers(1000, 30522, (256, 64)).astype(np.int64)).cuda()
mask = torch.ones((256, 64), dtype=torch.int64, device="cuda")
out = torch.empty((256, 384), device="cuda")
cs = torch.cuda.current_stream().cuda_stream
for b in ([int(os.environ['B'])] * 2 if os.environ.get('B') else (1, 1, 256, 256)):
    enc.encode_ids_device(ids.data_ptr(), mask.data_ptr(), b, 64, out.data_ptr(), cs)
torch.cuda.synchronize()
print("done")
