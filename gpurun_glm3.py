import sys, time, numpy as np, torch, ctypes
sys.path.insert(0, '/root/repo')
from pymc4_b200 import _cabi
L = _cabi.load_library()
n, P, C = 200000, 100, 1024
D, base = P + 2, 1
torch.manual_seed(0)
X = torch.randn(n, P, device="cuda"); y = (torch.rand(n, device="cuda") < 0.5).float()
ldxt = (n + 3) // 4 * 4
Xt = torch.zeros(P, ldxt, device="cuda"); Xt[:, :n] = X.T
theta = torch.randn(C, D, device="cuda") * 0.3
lp = torch.zeros(C, device="cuda"); grad = torch.zeros(C, D, device="cuda")
work = torch.empty(L.pm4_glm_work_bytes(n, P, C), dtype=torch.uint8, device="cuda")
err = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(3):
    rc = L.pm4_glm_logp_grad(X.data_ptr(), Xt.data_ptr(), ldxt, y.data_ptr(), n, P, theta.data_ptr(), D, base, C, lp.data_ptr(), grad.data_ptr(), D, work.data_ptr(), err.data_ptr(), None)
torch.cuda.synchronize()
print("rc", rc, int(err.item()))
