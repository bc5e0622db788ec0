import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from fake_spectra_b200 import fluxstatistics as fstat
rng=np.random.default_rng(0)
tau=torch.from_numpy(np.exp(rng.normal(-1,1.3,(8192,4460)))).cuda()
for _ in range(3):
    k,p=fstat.flux_power(tau, 4460.0)
torch.cuda.synchronize()
print(p[:3])
