import sys; sys.path.insert(0,'/root/repo')
import torch, fle_b200
from fle_b200 import GE
ctx = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(9)
m = 1 << 28
x = torch.randint(0, 1 << 20, (m,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
pb = torch.empty((m + 63) // 64, dtype=torch.int64, device="cuda"); cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(4): ctx.plain_eval_dev(x, GE, [1 << 19], bitmap=pb, count=cnt)
torch.cuda.synchronize(); print(int(cnt.item()))
