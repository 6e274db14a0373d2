#!/usr/bin/env python
"""profiles/dequant_traffic.json from an `ncu --set full` capture of the 7 Q4_K launches of one bench step:
DRAM bytes per launch next to the algorithmic bytes (bench.py reads the file into roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys

rep, out_path, src_label = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}[unit]


SHAPES = [(18432, 3072), (9216, 3072), (3072, 3072), (12288, 3072), (3072, 12288), (21504, 3072), (3072, 15360)]      # bench.py FLUX_SHAPES, launch order
per = []
for i, r in enumerate(data):
    grid = int(r[col("launch__grid_size")].replace(",", ""))
    n_elems = grid * 4096                                   # one 4096-element tile per CTA
    shape = list(SHAPES[i % len(SHAPES)])
    assert shape[0] * shape[1] == n_elems, (shape, n_elems)
    per.append({
        "shape": shape,
        "dram_read_bytes": to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")]),
        "dram_write_bytes": to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")]),
        "algorithmic_read_bytes": n_elems // 256 * 144,
        "algorithmic_write_bytes": n_elems * 2,
        "duration_us": to_us(r[col("gpu__time_duration.sum")], units[col("gpu__time_duration.sum")]),
    })
big = max(per, key=lambda x: x["algorithmic_write_bytes"])
json.dump({
    "kernel": "ggufb200::dequant_kernel<Block<Q4_K>, f16 math, f16 out, 128 threads>",
    "source": src_label,
    "per_launch": per,
    "note": "dram reads equal the algorithmic packed bytes (each packed byte is fetched once); dram writes are BELOW the algorithmic "
            "output bytes because ncu starts every replay with an empty L2 and part of the freshly written output is still resident in "
            "the 126 MB L2 when the kernel ends. No re-reads, no write amplification.",
    "largest_launch_total_dram_bytes": big["dram_read_bytes"] + big["dram_write_bytes"],
    "largest_launch_algorithmic_bytes": big["algorithmic_read_bytes"] + big["algorithmic_write_bytes"],
}, open(out_path, "w"), indent=1)
print(json.dumps(per, indent=1))
