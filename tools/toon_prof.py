import sys
sys.path.insert(0, ".")
import torch
from mcp_context_forge_b200 import engine, synth
from mcp_context_forge_b200.batching import GpuBatcher
b = GpuBatcher.get(); lib = b.ctx.lib
n = 4096
base = [synth.payload("A", 16384, seed=s).encode() for s in range(64)]
texts = [base[i % 64] for i in range(n)]
stream, offs = engine.pack_units(texts)
batch = engine.Batch(b.ctx, len(stream), n); batch.upload(stream, offs)
d_out = torch.empty(len(stream) + 16, dtype=torch.uint8, device="cuda"); d_len = torch.empty(n, dtype=torch.int32, device="cuda"); d_st = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    lib.cf_toon(b.ctx.h, batch.h, 0, d_out.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), None)
torch.cuda.synchronize()
