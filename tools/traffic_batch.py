"""One pass of the five plans at batch BATCH (environment, default 32: what the scene loop runs) for ncu: DRAM traffic of
every conv_umma launch of a batch.
usage: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_umma --csv --log-file out.csv python tools/traffic_batch.py
       python tools/traffic_batch.py --summarise out.csv profiles/rX_conv_umma_traffic.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) >= 4 and sys.argv[1] == "--summarise":
    import csv
    rows = list(csv.reader(open(sys.argv[2])))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    mn, mv, mu, idc = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, ids = 0.0, set()
    for r in rows[hi + 1:]:
        if len(r) > mv and r[mn].startswith("dram__bytes"):
            tot += float(r[mv].replace(",", "")) * scale[r[mu]]
            ids.add(r[idc])
    out = {"kernel": "conv_umma_kernel", "launches": len(ids), "dram_bytes_total": tot,
           "dram_bytes_per_launch": tot / max(1, len(ids)),
           "batch": int(os.environ.get("BATCH", "32")),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_umma over one pass of the five plans at this batch"}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(out)
    sys.exit(0)

from building_detection_b200.predict_model import CTORS, MODEL_NAMES  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
B = int(os.environ.get("BATCH", "32"))
x = torch.from_numpy((np.random.default_rng(0).integers(0, 256, (B, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)).cuda()
for name in MODEL_NAMES:
    nat = CTORS[name]().native_plan(B)
    nat.run_device(x.data_ptr(), 0, 0)
import torch  # noqa: E402
torch.cuda.synchronize()
