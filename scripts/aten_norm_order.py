"""How does ATen associate the sum of squares in torch.linalg.vector_norm / F.normalize over a contiguous [P,4] tensor?
Enumerates every order (sequential and pairwise, with and without fused multiply-add) and counts mismatching bits against
torch on 10^6 random quaternions. Result on torch 2.11 / B200: (x^2 + z^2) + (y^2 + w^2), products rounded separately -> 0
mismatches; every other order 7-19 %. csrc/gsr_math.cuh:quat_norm is spelled accordingly (fused-activation entry)."""
import itertools, torch
torch.manual_seed(0)
q = (torch.randn(1_000_000, 4, device="cuda") * 2.0)
n_t = torch.linalg.vector_norm(q, dim=1)
n_t2 = torch.nn.functional.normalize(q)
d = q.double()
def rnd(x): return x.float().double()
sq = [rnd(d[:, k] * d[:, k]) for k in range(4)]
def fma(k, c): return rnd(d[:, k] * d[:, k] + c)
res = {}
for perm in itertools.permutations(range(4)):
    a, b, c, e = perm
    res["seq%s" % (perm,)] = rnd(rnd(rnd(sq[a] + sq[b]) + sq[c]) + sq[e])
    res["seqfma%s" % (perm,)] = fma(e, fma(c, fma(b, sq[a])))
for (a, b), (c, e) in [((0, 1), (2, 3)), ((0, 2), (1, 3)), ((0, 3), (1, 2))]:
    res["pair%d%d_%d%d" % (a, b, c, e)] = rnd(rnd(sq[a] + sq[b]) + rnd(sq[c] + sq[e]))
    for x, y in [(a, b), (b, a)]:
        for z, w in [(c, e), (e, c)]:
            res["pairfma%d%d_%d%d" % (x, y, z, w)] = rnd(fma(y, sq[x]) + fma(w, sq[z]))
best = sorted(((int((torch.sqrt(v.float()) != n_t).sum()), k) for k, v in res.items()))[:6]
print(best)
# exact double accumulate then round (e.g. accumulation in higher precision)
ex = torch.sqrt((d * d).sum(1)).float()
print("double-acc:", int((ex != n_t).sum()), " float sum then sqrt in double:", int((torch.sqrt(rnd(rnd(rnd(sq[0]+sq[1])+sq[2])+sq[3])).float() != n_t).sum()))
