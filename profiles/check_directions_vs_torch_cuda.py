import os, sys
ROOT="/root/repo" if os.path.exists("/root/repo/oracle") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"oracle"))
import numpy as np, torch
import texir_oracle as O
import texir_code_b200 as tx
dev=torch.device("cuda:0")
rng=np.random.default_rng(1)
P=4096
nrm=rng.normal(size=(P,3)).astype(np.float32); nrm/=np.linalg.norm(nrm,axis=1,keepdims=True)
rough=rng.uniform(0.01,0.8,size=(P,1)).astype(np.float32)
torch.manual_seed(5); sh=torch.rand(P,1,2)
for mode in ("uniform","cosine","importance"):
    for N in (64, 2048):
        a=tx.generate_dir(torch.from_numpy(nrm).to(dev), N, None, mode, torch.from_numpy(rough).to(dev), shift=sh.to(dev))
        b=O.generate_dir(torch.from_numpy(nrm).to(dev), N, None, mode, torch.from_numpy(rough).to(dev), shift=sh.to(dev))
        c=O.generate_dir(torch.from_numpy(nrm), N, None, mode, torch.from_numpy(rough), shift=sh)
        print(mode, N, "ours vs torch-CUDA: mismatch frac %.5f maxabs %.2e | ours vs torch-CPU: %.5f | torch-CUDA vs torch-CPU %.5f"%(
            float((a!=b).any(-1).float().mean()), float((a-b).abs().max()), float((a.cpu()!=c).any(-1).float().mean()), float((b.cpu()!=c).any(-1).float().mean())))
