"""Small driver for profiling one incremental update (run under ncu): base of N accounts, M dirty."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import random_keys_torch, be_sort_key, splitmix64_torch
from reth_b200 import Engine, ResidentTrie

n, m = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
eng = Engine(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.use_torch_stream()
keys = random_keys_torch(5, n, dev)
keys = keys[torch.sort(be_sort_key(keys), stable=True).indices].contiguous()
accts = torch.zeros((n, 72), dtype=torch.uint8, device=dev)
accts[:, 32:40] = splitmix64_torch(9, n, dev).view(torch.uint8).view(n, 8)
root = torch.zeros(32, dtype=torch.uint8, device=dev)
trie = ResidentTrie.create_dev(eng, keys.view(torch.uint8).view(-1), accts.view(-1), None, n, root)
g = torch.Generator(device=dev); g.manual_seed(1)
for it in range(4):
    idx = torch.unique(torch.randint(0, n, (m + m // 8,), generator=g, device=dev))[:m]
    dk = keys[idx].contiguous().view(torch.uint8).view(-1)
    da = accts[idx].clone(); da[:, 0] = it + 1; da = da.view(-1)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push(f"update{it}")
    trie.update_dev(dk, da, None, int(idx.numel()), root)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done", eng.last_stats())
