"""A few eager G+D training steps and nothing else: target for `ncu --metrics gpu__time_duration.sum` launch lists.

    python profiles/run_step.py [N] [B] [steps] [gapt|mpgan]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mpgan_b200 import ops, presets, train

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
gapt = len(sys.argv) > 4 and sys.argv[4] == "gapt"
ops.set_precision(1)
torch.manual_seed(4)
dev = "cuda"
if gapt:
    G, D = presets.gapt_generator(num_hits=N).to(dev), presets.gapt_discriminator(num_hits=N).to(dev)
else:
    G, D = presets.mp_generator(num_hits=N).to(dev), presets.mp_discriminator(num_hits=N).to(dev)
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    G.load_state_dict(torch.load(os.path.join(gold, "mp_g_weights.pt"), map_location=dev))
    D.load_state_dict(torch.load(os.path.join(gold, "mp_d_seed4_weights.pt"), map_location=dev))
tr = train.GANTrainer(G, D, num_particles=N, latent_node_size=64 if gapt else 32)
data, labels, _ = train.synthetic_jets(B, N, dev, torch.Generator(device=dev).manual_seed(4))
for _ in range(steps):
    tr.step(data, labels)
torch.cuda.synchronize()
print("done")
