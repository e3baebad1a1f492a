"""One forward / data-gradient / weight-gradient launch of K7 at a given layer shape (for ncu)."""
import sys, torch
sys.path.insert(0, ".")
from maskunet_b200 import ops
cin, cout, hw = (int(a) for a in sys.argv[1:4])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
x = torch.randn(B, cin, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = torch.randn(cout, cin, 3, 3, device="cuda")
dy = torch.randn(B, cout, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
wf, wd = ops.conv_prep_weights(w, True)
for _ in range(2):
    ops.conv3x3_fwd(x, wf, True)
    ops.conv3x3_bwd_data(dy, wd)
    ops.conv3x3_bwd_weight(x, dy)
torch.cuda.synchronize()
