"""Per-command device times of the RAFT programs (encoders, one update iteration, mask + upsampling)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as m  # noqa: E402
from oracle import raft_oracle as RO  # noqa: E402

H, W, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tv = RO.build(0)
mine = m.raft_large()
mine.load_state_dict(tv.state_dict())
mine = mine.cuda().eval()
a, b = RO.frames(B, H, W)
with torch.no_grad():
    for _ in range(2):
        mine(a.cuda(), b.cuda())
P = mine._plans[(B, H, W)]
for name in ("E", "U", "M"):
    for _ in range(2):
        rows = P[name].run_timed()
    tot = sum(r["ms"] for r in rows)
    print("== program %s: %d commands, %.3f ms (sum of per-command times)" % (name, len(rows), tot))
    for r in rows:
        if r["ms"] * 1e3 >= 8.0:
            print("   %7.1f us  %-10s %s" % (r["ms"] * 1e3, r["kind"], r["name"]))
