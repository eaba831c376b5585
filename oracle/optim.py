"""optim.rmsprop / optim.adam oracle (float32), upstream `optim` rock defaults.

TEST INFRASTRUCTURE.  Call sites: /root/reference/src/lua/main.lua:135-144 (state = {learningRate = mp.lr}),
main.lua:301-302 (in-place update of the flat vector), main.lua:382-386 (lr decay).  The optimiser sources are NOT
in /root/reference (un-vendored, unpinned rock); the update rules are restated from the published algorithm:
  rmsprop: alpha=0.99, eps=1e-8;  m = alpha*m + (1-alpha)*g*g;  x -= lr * g / (sqrt(m) + eps);  m0 = 0
           (later optim versions initialise m to ones -> `m_init` flag)
  adam:    b1=0.9, b2=0.999, eps=1e-8; t+=1; m=b1 m+(1-b1) g; v=b2 v+(1-b2) g*g;
           x -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
"""
import numpy as np

F32 = np.float32


def rmsprop_init(n, m_init=0.0):
    return {"m": np.full(n, m_init, F32)}


def rmsprop_step(x, g, state, lr, alpha=0.99, eps=1e-8):
    m = state["m"]
    m[...] = (F32(alpha) * m + F32(1 - alpha) * (g * g)).astype(F32)
    x[...] = (x - F32(lr) * g / (np.sqrt(m) + F32(eps))).astype(F32)
    return x


def adam_init(n):
    return {"t": 0, "m": np.zeros(n, F32), "v": np.zeros(n, F32)}


def adam_step(x, g, state, lr, b1=0.9, b2=0.999, eps=1e-8):
    state["t"] += 1
    t = state["t"]
    m, v = state["m"], state["v"]
    m[...] = (F32(b1) * m + F32(1 - b1) * g).astype(F32)
    v[...] = (F32(b2) * v + F32(1 - b2) * (g * g)).astype(F32)
    step = F32(lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    x[...] = (x - step * m / (np.sqrt(v) + F32(eps))).astype(F32)
    return x


def lr_schedule(lr0, t, start_iter=1, decay=0.99, every=2500, after=50000):
    """learning rate in force DURING iteration t (1-based): main.lua:382-386 multiplies after iteration t when
    t >= lrdecayafter and (t-start_iter+1) % lrdecay_every == 0."""
    n = 0
    for tt in range(start_iter, t):
        if tt >= after and (tt - start_iter + 1) % every == 0:
            n += 1
    return lr0 * decay ** n
