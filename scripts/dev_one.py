import sys
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import torch
from dev_check import run
a = sys.argv[1:]
run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), a[4])
