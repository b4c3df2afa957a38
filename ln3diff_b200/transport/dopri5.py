"""Adaptive Dormand-Prince 5(4) solver: the `sampling_method='dopri5'` default of the reference's
`Sampler.sample_ode` (transport/transport.py:374-421 -> transport/integrators.py:112-119 ->
`torchdiffeq.odeint(fn, x, t, method='dopri5', atol=[1e-6], rtol=[1e-3])`).

torchdiffeq (0.2.3, environment_ln3diff.yml:297) is an un-vendored third-party dependency that is absent from
the image, so this file restates its PUBLISHED algorithm -- "parity unpinned" (SURVEY.md 8c): the accepted-step
sequence is data dependent and there is no copy of torchdiffeq here to pin it against; tests check the
solver against closed-form ODE solutions and its documented controller behaviour.  Restated pieces:
  * Dormand-Prince tableau (alpha, beta, c_sol = last beta row (FSAL), c_error = c_sol - 4th-order weights);
  * initial step selection (Hairer, Norsett, Wanner, "Solving ODEs I", II.4) with order 4;
  * error ratio = rms( err / (atol + rtol * max(|y0|, |y1|)) ), accept iff <= 1;
  * step controller dt' = dt * min(ifactor=10, max(safety=0.9 / ratio^(1/5), dfactor)), dfactor = 0.2 on a rejected
    step and 1 on an accepted one (the step never shrinks after an accept), ratio == 0 -> dt * 10;
  * dense output: quartic fit through (y0, f0, y_mid, y1, f1) with the DPS_C_MID mid-point weights; the outputs at
    the caller's time grid are interpolated, the solver never steps to them exactly.
Time-like quantities are float64 on the host (torchdiffeq keeps them in float64 tensors); the state keeps its dtype.
One error-norm reduction per attempted step is read back by the host (the accept / reject branch is host control
flow in torchdiffeq as well).  Every function evaluation is one `forward_with_cfg` = one CUDA-graph replay.
"""
from __future__ import annotations

import torch as th

_ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_BETA = ((1 / 5,),
         (3 / 40, 9 / 40),
         (44 / 45, -56 / 15, 32 / 9),
         (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
         (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
         (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
_C_ERROR = (35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
            -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0)
_C_MID = (6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)


def _rms(t: th.Tensor) -> float:
    return float(t.float().pow(2).mean().sqrt())


def _lincomb(y0, ks, coefs, dt):
    out = y0
    for k, c in zip(ks, coefs):
        if c != 0.0:
            out = out + k * (c * dt)
    return out


def _initial_step(fn, t0, y0, f0, order, rtol, atol):
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = fn(t0 + h0, y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return min(100 * h0, h1)


def _interp_eval(coef, t0, t1, t):
    x = (t - t0) / (t1 - t0)
    total = coef[0] + x * coef[1]
    xp = x
    for c in coef[2:]:
        xp = xp * x
        total = total + xp * c
    return total


def odeint_dopri5(fn, y0: th.Tensor, t, rtol: float = 1e-3, atol: float = 1e-6, first_step: float | None = None,
                  safety: float = 0.9, ifactor: float = 10.0, dfactor: float = 0.2, max_num_steps: int = 2 ** 31 - 1,
                  stats: dict | None = None) -> th.Tensor:
    """fn(t: float, y) -> dy/dt; `t` an increasing 1-D grid (tensor or sequence); returns the stacked solution
    (len(t), *y0.shape) with solution[0] = y0.  `stats` (optional dict) receives nfe / accepted / rejected."""
    ts = [float(v) for v in (t.tolist() if isinstance(t, th.Tensor) else t)]
    assert all(b > a for a, b in zip(ts, ts[1:])), "t must be strictly increasing"
    nfe = [0]

    def f(tt, yy):
        nfe[0] += 1
        return fn(tt, yy)

    f0 = f(ts[0], y0)
    dt = first_step if first_step is not None else _initial_step(f, ts[0], y0, f0, 4, rtol, atol)
    y, t0, t1 = y0, ts[0], ts[0]
    interp = [y0] * 5
    sol = [y0]
    accepted = rejected = 0
    for t_out in ts[1:]:
        n = 0
        while t_out > t1:
            assert n < max_num_steps, f"max_num_steps exceeded ({n}>={max_num_steps})"
            assert t1 + dt > t1, f"underflow in dt {dt}"
            ks = [f0]
            for alpha, beta in zip(_ALPHA, _BETA):
                yi = _lincomb(y, ks, beta, dt)
                ks.append(f(t1 + alpha * dt, yi))
            y1, f1 = yi, ks[-1]                                     # FSAL: the last stage is the solution
            err = _lincomb(th.zeros_like(y), ks, _C_ERROR, dt)
            tol = atol + rtol * th.max(y.abs(), y1.abs())
            ratio = _rms(err / tol)
            accept = ratio <= 1.0
            if accept:
                y_mid = _lincomb(y, ks, _C_MID, dt)
                a = 2 * dt * (f1 - f0) - 8 * (y1 + y) + 16 * y_mid
                b = dt * (5 * f0 - 3 * f1) + 18 * y + 14 * y1 - 32 * y_mid
                c = dt * (f1 - 4 * f0) - 11 * y - 5 * y1 + 16 * y_mid
                interp = [y, dt * f0, c, b, a]
                t0, t1, y, f0 = t1, t1 + dt, y1, f1
                accepted += 1
            else:
                rejected += 1
            if ratio == 0:
                dt = dt * ifactor
            else:
                df = 1.0 if ratio < 1 else dfactor
                dt = dt * min(ifactor, max(safety / ratio ** 0.2, df))
            n += 1
        sol.append(_interp_eval(interp, t0, t1, t_out))
    if stats is not None:
        stats.update(nfe=nfe[0], accepted=accepted, rejected=rejected)
    return th.stack(sol, 0)
