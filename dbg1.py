import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from tests import parity as P
from oracle import cvae_gan_oracle as O
B, seed = 64, 69
orc, eng, g = P.make_pair(10, 5, B, seed=seed)
x, y = P.make_data(10, 5, [B] * 5, seed=1)
xb = x[y == 4][:B].contiguous()
eng.zero_grads()
ref, got, grads = P.run_step("g", orc, eng, xb, 4, g, lambda_class=0.25, update=False)
sd = orc.sd["encoder"]
with torch.no_grad():
    h = torch.cat([xb, F.one_hot(torch.full([B],4),5).float()],1)
    hs, ys = [], []
    for li in (0,3,6):
        h = F.linear(h, sd[f"encoder.{li}.weight"], sd[f"encoder.{li}.bias"])
        hs.append(h)
        pre = F.batch_norm(h, None, None, sd[f"encoder.{li+1}.weight"], sd[f"encoder.{li+1}.bias"], True, 0.1, 1e-5)
        ys.append(pre)
        h = F.leaky_relu(pre, 0.2)
for i in range(3):
    mine = eng.debug_read(f"e_h{i}", B).cpu()
    print(f"e_h{i} max abs err", float((mine-hs[i]).abs().max()), "scale", float(hs[i].abs().max()))
print("oracle y2[7,57] =", float(ys[1][7,57]), " h2[7,57] oracle", float(hs[1][7,57]), "mine", float(eng.debug_read("e_h1",B)[7,57]))
h2m = eng.debug_read("e_h1", B).cpu().double()
mean = h2m.mean(0); var = h2m.var(0, unbiased=False)
gam = sd["encoder.4.weight"].detach().double(); bet = sd["encoder.4.bias"].detach().double()
ym = (h2m-mean)/torch.sqrt(var+1e-5)*gam+bet
print("my-h-based y2[7,57] (double) =", float(ym[7,57]))
dy = eng.debug_read("e_dy1", B).cpu()
print("e_dy1[:,57] mine:", dy[:10,57].tolist())
eng.close()
