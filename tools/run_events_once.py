"""The event encoders once at BASELINE configs[2] (10 M events, 1280x720), for ncu captures."""
import sys, torch
sys.path.insert(0, "/root/repo")
from ebfi_be_b200 import encodings
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(7)
NEV, EH, EW = 10_000_000, 720, 1280
xs = torch.randint(0, EW, (NEV,), generator=g).float().to(dev)
ys = torch.randint(0, EH, (NEV,), generator=g).float().to(dev)
ts = torch.sort(torch.rand(NEV, generator=g))[0].to(dev)
ps = (torch.randint(0, 2, (NEV,), generator=g) * 2 - 1).float().to(dev)
rx, ry, rt, rp = xs.to(torch.int16), ys.to(torch.int16), ts.double() * 0.5 + 100.0, ps.to(torch.int8)
for _ in range(3):
    encodings.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(EH, EW))
    encodings.events_to_stack(xs, ys, ts, ps, 16, sensor_size=(EH, EW))
    encodings.events_raw_to_stack(rx, ry, rt, rp, 16, (EH, EW))
torch.cuda.synchronize()
