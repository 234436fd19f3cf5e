"""ncu CSV (raw page, one row per captured launch) -> profiles/traffic_cfg3_tsp65.json / _cold.json with the kernel-source hash.
    python scripts/traffic_from_ncu.py gpurun_out/r02_traffic.csv 100 20"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
it_steady, it_cold = int(sys.argv[2]), int(sys.argv[3])


def metric(launch_id, name):
  for r in rows:
    if r["ID"] == str(launch_id) and r["Metric Name"] == name:
      v = float(r["Metric Value"].replace(",", ""))
      u = r["Metric Unit"].lower()
      scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
               "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(u, 1.0)
      return v * scale
  raise KeyError((launch_id, name))


ids = sorted({int(r["ID"]) for r in rows})
# the two long launches are the windows (tables launches are microseconds)
dur = {i: metric(i, "gpu__time_duration.sum") for i in ids}
big = sorted(ids, key=lambda i: -dur[i])
print({i: round(dur[i] * 1e3, 3) for i in ids})
for tag, iters, want in (("cfg3_tsp65", it_steady, None), ("cfg3_tsp65_cold", it_cold, None)):
  pass
steady_id, cold_id = int(sys.argv[4]), int(sys.argv[5])
for tag, iters, lid in (("cfg3_tsp65", it_steady, steady_id), ("cfg3_tsp65_cold", it_cold, cold_id)):
  by = metric(lid, "dram__bytes_read.sum") + metric(lid, "dram__bytes_write.sum")
  out = {"workload": tag, "src_hash": bench.src_hash(), "iters_in_launch": iters, "dram_bytes_per_launch": by, "dram_bytes_per_iter": by / iters,
         "dram_read_bytes": metric(lid, "dram__bytes_read.sum"), "dram_write_bytes": metric(lid, "dram__bytes_write.sum"),
         "gpu_time_s_under_ncu": dur[lid], "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
         "-k regex:pdhg_coop python scripts/traffic_capture.py %d (launch id %d); includes the launch's one-off state copy-in / copy-out" % (it_steady, lid)}
  json.dump(out, open(os.path.join(ROOT, "profiles", "traffic_%s.json" % tag), "w"), indent=1)
  print(tag, "DRAM bytes / iteration: %.1f MB" % (by / iters / 1e6))
