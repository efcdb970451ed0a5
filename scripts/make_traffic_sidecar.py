"""profiles/r02_traffic.json from the `ncu --set full` captures of scripts/gemm_prof.py (first profiled launch = ViT fc1 3202x4096x1024):
dram__bytes_read.sum + dram__bytes_write.sum of that launch, read by bench.py as `roofline.traffic`."""
import csv, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for key, rep, name in (("tcgen05-split", "gpurun_out/ncu_gemm_split_final.ncu-rep", "split-fp16 pairs"), ("tcgen05", "gpurun_out/ncu_gemm_fp16_final.ncu-rep", "fp16")):
    path = os.path.join(root, rep)
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    def val(metric):
        i = hdr.index(metric)
        v = float(first[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    out[key] = {"dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                "launch": f"vit fc1 3202x4096x1024, {name}, {first[hdr.index('Kernel Name')][:60]} (ncu --set full, profiles/r02_ncu_gemm_fc1_qkv_{key.replace('-', '_')}_final.txt)",
                "algorithmic_bytes": 3202 * 1024 * (8 if "split" in key else 2) // (2 if "split" in key else 1) * (2 if "split" in key else 1) + 4096 * 1024 * (4 if "split" in key else 2) + 3202 * 4096 * (4 if "split" in key else 2)}
json.dump(out, open(os.path.join(root, "profiles", "r02_traffic.json"), "w"), indent=1)
print(out)
