"""CPU emulation of the float32 restricted chains (K2): which rounding is responsible for the error that
survives the sum over pairs?  FMA emulated exactly through float64."""
import sys, numpy as np
sys.path.insert(0, '.')
rng = np.random.default_rng(1)
NP_, n, T, K, P, D = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 5, 50, 65, 49, 64
cent = 10 * rng.standard_normal((K, D))
con = rng.integers(0, K, (NP_, n))
v = (cent[con] + rng.standard_normal((NP_, n, D))).astype(np.float32).astype(np.float64)
W = 0.01 * rng.standard_normal((K, D)) * np.sqrt(512 / D)
lg = v @ W.T
lg -= lg.max(-1, keepdims=True)
pz = np.exp(lg); pz /= pz.sum(-1, keepdims=True)            # (NP, n, K)
pw = 1.0 / np.arange(1, P + 1) ** 1.2
x = rng.choice(P, size=(NP_, T), p=pw / pw.sum())
obs = rng.random((K, P)); obs /= obs.sum(1, keepdims=True)  # like the reference's random init
if len(sys.argv) > 2 and sys.argv[2] == 'peaked':
    obs = rng.random((K, P)) ** 8; obs /= obs.sum(1, keepdims=True)
A = rng.random((n, n)) + 0.5; A /= A.sum(1, keepdims=True)
pi = np.full(n, 1.0 / n)

f32 = lambda a: np.asarray(a, np.float64).astype(np.float32).astype(np.float64)
def split(a):
    hi = f32(a); return hi, f32(a - hi)

def chains(mode):
    """mode: dict(prec='f64'|'f32', o_lo, e_lo, A_lo).  Returns cC (NP, n, K)."""
    r = (lambda a: a) if mode['prec'] == 'f64' else f32
    ot = obs.T[x]                                 # (NP, T, K)  o_k(x_t)
    e = np.einsum('pjk,ptk->ptj', pz, ot)         # (NP, T, n) marginal emissions (float64, as the kernel)
    if mode['prec'] == 'f64':
        o_hi, o_lo, e_hi, e_lo, A_hi, A_lo = ot, 0 * ot, e, 0 * e, A, 0 * A
    else:
        o_hi, o_lo = split(ot); e_hi, e_lo = split(e); A_hi, A_lo = split(A)
        if not mode.get('o_lo'): o_lo = 0 * o_lo
        if not mode.get('e_lo'): e_lo = 0 * e_lo
        if not mode.get('A_lo'): A_lo = 0 * A_lo
    pi_ = r(pi)
    eye = np.eye(n, dtype=bool)                   # [i, j]
    def emis(t):
        # (NP, i, k, j): hi and lo parts
        eh = np.broadcast_to(e_hi[:, t][:, None, None, :], (NP_, n, K, n)).copy()
        el = np.broadcast_to(e_lo[:, t][:, None, None, :], (NP_, n, K, n)).copy()
        oh = np.broadcast_to(o_hi[:, t][:, None, :, None], (NP_, n, K, n))
        ol = np.broadcast_to(o_lo[:, t][:, None, :, None], (NP_, n, K, n))
        m = np.broadcast_to(eye[None, :, None, :], (NP_, n, K, n))
        eh[m] = oh[m]; el[m] = ol[m]
        return eh, el
    eh, el = emis(0)
    F = r(r(pi_ * eh) + pi_ * el) if mode['prec'] != 'f64' else pi_ * eh
    for t in range(1, T):
        acc = np.zeros_like(F)
        for l in range(n):
            acc = r(acc + F[..., l:l + 1] * A_hi[l][None, None, None, :])
        if mode.get('A_lo'):
            for l in range(n):
                acc = r(acc + F[..., l:l + 1] * A_lo[l][None, None, None, :])
        eh, el = emis(t)
        if mode['prec'] == 'f64': F = acc * eh
        elif mode.get('o_lo') or mode.get('e_lo'): F = r(acc * eh + r(acc * el))     # fma(acc, hi, acc*lo)
        else: F = r(acc * eh)
        if mode['prec'] != 'f64' and t % 8 == 0:
            m = F.max(-1, keepdims=True); sc = 2.0 ** -np.floor(np.log2(np.maximum(m, 1e-300))); F = F * sc; 
            scale = scale + np.log2(sc[..., 0]) if 'scale' in dir() else np.log2(sc[..., 0])
    L = F.sum(-1)                                 # (NP, n, K)
    if mode['prec'] != 'f64' and 'scale' in dir():
        L = L * 2.0 ** (-(scale - scale.max(-1, keepdims=True)))
    num = pz * L
    return num / num.sum(-1, keepdims=True)

c64 = chains(dict(prec='f64'))
G64 = np.einsum('pik,pid->kd', c64 - pz, v)
print('pairs %d  |G|max %.3e' % (NP_, np.abs(G64).max()))
for name, mode in [('f32 plain', dict(prec='f32')), ('f32 o_lo', dict(prec='f32', o_lo=1)),
                   ('f32 o_lo e_lo', dict(prec='f32', o_lo=1, e_lo=1)),
                   ('f32 o_lo e_lo A_lo', dict(prec='f32', o_lo=1, e_lo=1, A_lo=1))]:
    c = chains(mode)
    d = c - c64
    G = np.einsum('pik,pid->kd', d, v)
    print('%-22s cC max abs %.2e  rms %.2e | grad err / max|G| %.2e   (random-walk expectation %.2e)' % (
        name, np.abs(d).max(), np.sqrt((d ** 2).mean()), np.abs(G).max() / np.abs(G64).max(),
        np.sqrt((d ** 2).mean()) * np.sqrt(NP_ * n) * np.abs(v).mean() / np.abs(G64).max()))
