"""profiles/traffic.json from ncu launch lists: average DRAM bytes (read + write) per launch of the three stage kernels
(k_pipe<1|2|3> = the 'k_euler_stage' of bench.py's roofline).  usage: traffic_from_launch_list.py ne120_q35=profiles/x.csv ..."""
import csv, json, os, sys
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, path = arg.split("=")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
    h = rows[0]
    ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, n = 0.0, 0
    for r in rows[1:]:
        if not any(("k_pipe<%d>" % op) in r[ik] for op in (1, 2, 3)):
            continue
        v = float(r[iv].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1)
        if r[im].startswith("dram__bytes"):
            tot += v
        if r[im] == "dram__bytes_read.sum":
            n += 1
    out[key] = {"k_euler_stage_dram_bytes_per_launch": tot / max(n, 1), "launches": n, "source": os.path.relpath(path, os.path.dirname(out_path))}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
