"""1M-jet generation sweep (BASELINE configs[2]; reference gen.py / train.gen_multi_batch, train.py:226-282) through
train.gen_multi_batch: batches of 4096 jets, output streamed into ONE pinned host buffer.  Prints end-to-end jets/s
(wall clock around the whole sweep, including every device->host copy).

    python profiles/gen_sweep.py [N] [num_jets] [batch]
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import ops, presets, train

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda", 0)
ops.set_precision(1)
torch.manual_seed(4)
G = presets.mp_generator(num_hits=N).to(dev).eval()
gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
G.load_state_dict(torch.load(os.path.join(gold, "mp_g_weights.pt"), map_location=dev))
g = torch.Generator().manual_seed(4)
n = torch.randint(1, N + 1, (total,), generator=g)
labels = (n.float() * torch.tensor(1.0 / N)).unsqueeze(1).pin_memory()
train.gen_multi_batch(G, 4 * batch, batch, N, labels=labels[:4 * batch])   # warm-up (allocator, first launches)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = train.gen_multi_batch(G, total, batch, N, labels=labels)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ok = bool(torch.equal((out[:20000, :, 3] > 0).sum(1), n[:20000]))
print(f"N={N}: {total} jets in {dt:.3f} s = {total / dt:,.0f} jets/s end to end (batch {batch}, output "
      f"{out.numel() * 4 / 1e9:.2f} GB pinned host memory); particle counts of the first 20000 jets match labels: {ok}")
