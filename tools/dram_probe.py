"""How many DRAM bytes does ncu count for a plain streaming read of N bytes? (calibrates dram__bytes_read)."""
import torch
torch.cuda.set_device(0)
x = torch.randint(0, 255, (1800 * 3110400,), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    y = x.view(torch.int32).sum()
torch.cuda.synchronize()
print(int(y))
