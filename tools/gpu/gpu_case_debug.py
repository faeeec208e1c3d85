import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import golden_scenes as gs
import test_gpu_parity as tg
name = sys.argv[1]
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
bm, ids, post, rays, _ = tg.run_engine(name)
for k in range(4):
    print("ids[%d] mismatch" % k, (ids[..., k] != g["ids"][..., k]).sum())
bad = np.argwhere(ids[..., 1] != g["ids"][..., 1])
for y, x in bad[:: max(1, len(bad) // 10)][:10]:
    print((x, y), "eng", ids[y, x], post[y, x, :4], "gold", g["ids"][y, x], g["post"][y, x, :4])
d = np.abs(bm.astype(int) - g["bitmap"].astype(int)).max(-1)
print("rgb>2:", (d > 2).sum(), "of", d.size)
