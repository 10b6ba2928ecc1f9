"""Per-layer timing of the conv stack on the GPU (measurement helper, not part of the product).
    python tools/profile_layers.py [H W C NC batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
from yolo3_b200 import Engine, weights  # noqa: E402

a = [int(v) for v in sys.argv[1:]] + [512, 512, 1, 1, 32][len(sys.argv) - 1:]
H, W, C, NC, B = a[:5]
eng = Engine((H, W, C), NC, None, max_batch=B)
eng.load_weights(weights.random_init(C, NC, 3, seed=0, randomize_bn=True))
rep = eng.profile_layers(B, 5)
rows = [r.split(",") for r in rep.strip().split("\n")]
print("%-20s %-5s k s %5s %5s %4s %4s  patch   bn bk %6s %8s %7s %8s" % ("name", "kind", "cin", "cout", "oh", "ow", "tiles", "ms", "TF/s", "GB/s"))
tot = 0.0
for r in rows[1:]:
    tot += float(r[13])
    print("%-20s %-5s %s %s %5s %5s %4s %4s %3sx%-3s %3s %2s %6s %8.4f %7.1f %8.1f" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], float(r[13]), float(r[14]), float(r[15])))
print("total ms %.3f  (batch %d)" % (tot, B))
