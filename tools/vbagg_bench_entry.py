import sys, json, torch
sys.path.insert(0, "/root/repo")
import bench
print(json.dumps(bench.run_vbagg(torch.device("cuda", 0)), indent=1))
