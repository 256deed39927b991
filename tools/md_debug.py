import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
stream = torch.cuda.Stream(device=dev)
os.environ["MD_DEBUG"] = "1"
print(bench.section_md(dev, stream, 2))
