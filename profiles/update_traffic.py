"""profiles/traffic.json from an `ncu --set full` raw CSV of profiles/probe_kernels.py (capture.sh): DRAM bytes per launch
of the dominant kernel (score_select_tc_kernel, D = 8) at the C4 and C2 shapes, stamped with the digest of the kernel
source it was captured on — bench.py only reports `roofline.traffic` when that digest matches the build it runs.

    python profiles/update_traffic.py gpurun_out/r2l_kernels_raw.csv profiles/r2l_kernels_full.txt
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pivotcvae_b200 import build as b  # noqa: E402

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def val(row, name):
    i = hdr.index(name)
    return float(row[i].replace(",", "")) * scale.get(units[i], 1.0)


filt = [r for r in rows[2:] if "score_select_tc_kernel<1" in r[hdr.index("Kernel Name")] or
        ("score_select_tc_kernel(" in r[hdr.index("Kernel Name")])]
# probe order: select (C4: 1 M x 20480), then select_d128 (another instantiation), then select_c2 (50 k x 10240), ...
out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (score_select_tc_kernel<1>), one launch, "
                "from the ncu --set full capture summarised in " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]),
       "kernel_digest": b.kernel_digest("score_select_tc.cu")}
shapes = [("c4", 20480, 1000000, 33474560), ("c2", 10240, 50000, 2009600)]
for (name, M, N, alg), r in zip(shapes, filt[:2]):
    out[name] = {"kernel": "score_select_tc_kernel<1> M=%d N=%d" % (M, N), "M": M, "N": N,
                 "bytes": int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")), "algorithmic_bytes": alg,
                 "duration_ms": float(r[hdr.index("gpu__time_duration.sum")]) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
