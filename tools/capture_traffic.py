"""DRAM traffic of ONE launch of the bench kernel at the bench size (ncu, dram__bytes_read/write only), keyed to a hash of the kernel
sources so that bench.py reports it only while the kernel is unchanged.  Run on the GPU box:
    python tools/capture_traffic.py            -> gpurun_out/traffic_denoise.json   (copy to profiles/<round>/)"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["jeicyboodsp_b200/csrc/kernels_stream.cuh", "jeicyboodsp_b200/csrc/kernels_stft.cuh", "jeicyboodsp_b200/csrc/jdsp_device.cuh"]


def source_hash() -> str:
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()


if __name__ == "__main__":
    streams, seconds = 4096, 60.0
    out = os.path.join(ROOT, "gpurun_out", "traffic_denoise")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none", "-k", "regex:denoise_stream",
           "-c", "2", "--csv", "--log-file", out + ".csv", sys.executable, os.path.join(ROOT, "tools", "prof_denoise.py"),
           "--streams", str(streams), "--seconds", str(seconds), "--iters", "1"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    rows = [r for r in csv.reader(open(out + ".csv")) if len(r) > 5]
    hdr = rows[0]
    iname, imet, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = {}
    for r in rows[1:]:
        v = float(r[ival].replace(",", ""))
        unit = r[iunit]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(unit, 1)
        launches.setdefault((r[hdr.index("ID")], r[iname]), {})[r[imet]] = v * scale
    per = [{"kernel": k[1], **m} for k, m in launches.items()]
    n = int(seconds * 16000)
    nb = n // 256
    doc = {"streams": streams, "samples_per_stream": n, "algorithmic_bytes_per_launch": streams * (n + (nb - 2) * 256) * 2, "launches": per,
           "dram_bytes_per_launch": sum(p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"] for p in per) / len(per),
           "kernel_sources": KERNEL_SOURCES, "kernel_sources_sha256": source_hash(),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one SS and one Wiener launch of tools/prof_denoise.py at the bench size"}
    json.dump(doc, open(out + ".json", "w"), indent=1)
    print(json.dumps(doc)[:600])
