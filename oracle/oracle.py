"""ctypes front-end of the CPU oracle (oracle/sgmm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.  Each wrapper names the reference
lines its C counterpart restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsgmm_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "sgmm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libsgmm_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


_lib = None
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class _Env(C.Structure):
    _fields_ = [("phi", C.c_double), ("tick_size", C.c_double), ("fee_rate", C.c_double),
                ("inventory", C.c_int64), ("cash", C.c_double),
                ("i_max", C.c_int64), ("i_min", C.c_int64)]


class _Info(C.Structure):
    _fields_ = [("reward", C.c_double), ("pnl_reward", C.c_double),
                ("inventory_reward", C.c_double), ("fee_paid", C.c_double),
                ("fill_buy", C.c_int32), ("fill_sell", C.c_int32)]


class _Trace(C.Structure):
    _fields_ = [("off_a", _i32p), ("off_b", _i32p), ("adv_a", _i32p), ("adv_b", _i32p),
                ("fill_buy", _i32p), ("fill_sell", _i32p), ("inventory", _i32p),
                ("cash", _f64p), ("reward", _f64p), ("pnl_reward", _f64p),
                ("inventory_reward", _f64p), ("fee_paid", _f64p),
                ("raw_a", _f32p), ("raw_b", _f32p)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_genome_len.restype = C.c_int64
        L.oracle_genome_len.argtypes = [C.c_int]
        L.oracle_argmax.restype = C.c_int64
        L.oracle_argmax.argtypes = [_f64p, C.c_int64]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_env_init.argtypes = [C.POINTER(_Env), C.c_double, C.c_double, C.c_double]
        L.oracle_env_step.argtypes = [C.POINTER(_Env), _i64p, _i64p] + [C.c_double] * 5 + [C.POINTER(_Info)]
        L.oracle_mlp_forward.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.oracle_adv_forward.argtypes = [_f32p, _f32p, _f32p, _i64p]
        L.oracle_quantise.argtypes = [_f32p, _i64p]
        L.oracle_normal4.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _f32p]
        L.oracle_mutate.argtypes = [_f32p, C.c_int64, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64, _f32p]
        L.oracle_rollout.argtypes = [_f32p, _f32p, C.c_int, C.c_int64, _f32p, _f32p,
                                     _f64p, _f64p, _f64p, _f64p, _f64p,
                                     C.c_double, C.c_double, C.c_double, _i32p,
                                     _f64p, _i32p, C.POINTER(_Trace)]
        L.oracle_rollout_population.argtypes = [
            C.c_int64, C.c_int, _f32p, _f32p, C.c_float, _f32p, _f32p, C.c_float, C.c_int,
            C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, _f32p, _f32p,
            _f64p, _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double,
            _f64p, _i32p, C.c_int]
        _lib = L
    return _lib


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def genome_len(hidden: int = 32) -> int:
    return int(lib().oracle_genome_len(hidden))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def normalise(bundle, train_stats):
    """drl_engine.py:33-35: ``(s[t]-m)/s`` in the caller's numpy dtypes, then the float32 cast of
    ``torch.tensor(..., dtype=torch.float32)``.  Vectorised numpy applies the identical scalar
    expression element by element (same dtype promotion under NEP 50)."""
    s1, s2 = np.asarray(bundle[0]), np.asarray(bundle[1])
    z1 = ((s1 - train_stats['s1_m']) / train_stats['s1_s']).astype(np.float32)
    z2 = ((s2 - train_stats['s2_m']) / train_stats['s2_s']).astype(np.float32)
    return z1, z2


class Env:
    """FTPEnv restated (market_env.py:8-67)."""

    def __init__(self, phi=0.01, tick_size=0.01, fee_rate=0.0):
        self._e = _Env()
        lib().oracle_env_init(C.byref(self._e), phi, tick_size, fee_rate)

    inventory = property(lambda s: int(s._e.inventory), lambda s, v: setattr(s._e, "inventory", int(v)))
    cash = property(lambda s: float(s._e.cash), lambda s, v: setattr(s._e, "cash", float(v)))

    def reset(self):
        self._e.inventory = 0
        self._e.cash = 0.0
        return 0, 0.0

    def step(self, action, mid_next, best_ask, best_bid, buy_max, sell_min, adv_action=None):
        act = (C.c_int64 * 2)(int(action[0]), int(action[1]))
        adv = None
        if adv_action is not None:
            r = np.round(np.asarray(adv_action)).astype(np.int64)     # market_env.py:26
            adv = (C.c_int64 * 2)(int(r[0]), int(r[1]))
        info = _Info()
        lib().oracle_env_step(C.byref(self._e), act, adv, float(mid_next), float(best_ask),
                              float(best_bid), float(buy_max), float(sell_min), C.byref(info))
        return info.reward, {'pnl_reward': info.pnl_reward, 'inventory_reward': info.inventory_reward,
                             'fee_paid': info.fee_paid, 'fill_buy': info.fill_buy,
                             'fill_sell': info.fill_sell}


def mlp_forward(genome, x, hidden: int = 32):
    """TradingPolicy.forward (model.py:24-26) in SGMM-F32 order; returns raw float32[2]."""
    g = _f32(genome)
    xx = _f32(x)
    out = np.zeros(2, np.float32)
    lib().oracle_mlp_forward(_p(g, _f32p), hidden, _p(xx, _f32p), _p(out, _f32p))
    return out


def quantise(raw):
    r = _f32(raw)
    out = np.zeros(2, np.int64)
    lib().oracle_quantise(_p(r, _f32p), _p(out, _i64p))
    return out


def adv_forward(genome, x):
    g = _f32(genome)
    xx = _f32(x)
    pre = np.zeros(2, np.float32)
    d = np.zeros(2, np.int64)
    lib().oracle_adv_forward(_p(g, _f32p), _p(xx, _f32p), _p(pre, _f32p), _p(d, _i64p))
    return pre, d


def normal4(seed, generation, individual, block):
    out = np.zeros(4, np.float32)
    lib().oracle_normal4(seed, generation, individual, block, _p(out, _f32p))
    return out


def mutate(master, sigma, seed, generation, individual):
    m = _f32(master)
    out = np.empty_like(m)
    lib().oracle_mutate(_p(m, _f32p), m.size, sigma, seed, generation, individual, _p(out, _f32p))
    return out


def argmax(fitness) -> int:
    f = _f64(fitness)
    return int(lib().oracle_argmax(_p(f, _f64p), f.size))


def _unpack(bundle_z):
    z1, z2, mid, ask, bid, bmax, smin = bundle_z
    return _f32(z1), _f32(z2), _f64(mid), _f64(ask), _f64(bid), _f64(bmax), _f64(smin)


def rollout(mm_genome, adv_genome, bundle_z, phi, tick, fee, hidden=32, forced_actions=None,
            trace=False):
    """evaluate_individual (drl_engine.py:9-67).  ``bundle_z`` = (z1, z2, mid_next, best_ask,
    best_bid, buy_max, sell_min) with z from :func:`normalise`.  ``forced_actions`` int32[T,2]
    replaces the policy (teacher-forced replay).  Returns (fitness, trades[, trace dict])."""
    z1, z2, mid, ask, bid, bmax, smin = _unpack(bundle_z)
    T = z1.size
    g = _f32(mm_genome)
    a = _f32(adv_genome)
    fa = None if forced_actions is None else np.ascontiguousarray(forced_actions, np.int32)
    fit = C.c_double()
    tr = C.c_int32()
    tdict, tstruct = None, None
    if trace:
        tdict = {k: np.zeros(T, np.int32) for k in
                 ("off_a", "off_b", "adv_a", "adv_b", "fill_buy", "fill_sell", "inventory")}
        tdict.update({k: np.zeros(T, np.float64) for k in
                      ("cash", "reward", "pnl_reward", "inventory_reward", "fee_paid")})
        tdict.update({k: np.zeros(T, np.float32) for k in ("raw_a", "raw_b")})
        tstruct = _Trace(*[_p(tdict[n], t) for n, t in _Trace._fields_])
    lib().oracle_rollout(_p(g, _f32p), _p(a, _f32p), hidden, T, _p(z1, _f32p), _p(z2, _f32p),
                         _p(mid, _f64p), _p(ask, _f64p), _p(bid, _f64p), _p(bmax, _f64p),
                         _p(smin, _f64p), phi, tick, fee, _p(fa, _i32p), C.byref(fit), C.byref(tr),
                         None if tstruct is None else C.byref(tstruct))
    if trace:
        return fit.value, tr.value, tdict
    return fit.value, tr.value


def rollout_population(bundle_z, phi, tick, fee, *, genomes=None, master=None, sigma=0.05,
                       adv_genomes=None, adv_master=None, adv_sigma=0.05, use_adv=False,
                       seed=0, generation=0, first_index=0, count=None, hidden=32, nthreads=0):
    """The reference's Pool.starmap(evaluate_individual) (drl_engine.py:104-115) over P individuals."""
    z1, z2, mid, ask, bid, bmax, smin = _unpack(bundle_z)
    g = _f32(genomes)
    P = int(count if count is not None else g.shape[0])
    fit = np.zeros(P, np.float64)
    trd = np.zeros(P, np.int32)
    m, ag, am = _f32(master), _f32(adv_genomes), _f32(adv_master)
    lib().oracle_rollout_population(P, hidden, _p(g, _f32p), _p(m, _f32p), sigma, _p(ag, _f32p),
                                    _p(am, _f32p), adv_sigma, int(bool(use_adv)), seed, generation,
                                    first_index, z1.size, _p(z1, _f32p), _p(z2, _f32p),
                                    _p(mid, _f64p), _p(ask, _f64p), _p(bid, _f64p), _p(bmax, _f64p),
                                    _p(smin, _f64p), phi, tick, fee, _p(fit, _f64p), _p(trd, _i32p),
                                    nthreads)
    return fit, trd


def bundle_windows(ask1, bid1, p_buy_max, p_sell_min, step, n):
    """pipeline/agent_trainer.py:47-73 for one day: returns (mid_next, best_ask, best_bid, buy_max, sell_min)."""
    a, b, mx, mn = (_f64(x) for x in (ask1, bid1, p_buy_max, p_sell_min))
    m = max(int(n) - 1, 0)
    out = [np.empty(m, np.float64) for _ in range(5)]
    L = lib()
    L.oracle_bundle_windows.restype = C.c_int64
    r = L.oracle_bundle_windows(C.c_int64(len(a)), _p(a, _f64p), _p(b, _f64p), _p(mx, _f64p),
                                _p(mn, _f64p), C.c_int64(int(step)), C.c_int64(int(n)),
                                *[_p(o, _f64p) for o in out])
    if r < 0:
        raise ValueError("more signals than sampled events")
    return tuple(o[:r] for o in out)


ANALYTICS_KEYS = ("Total PnL", "MAP (Risk)", "PnLMAP (Eff)", "Max DD", "Sharpe", "Trades")


def analytics(wealth, inventory, is_trade):
    """analytics/mm_analyzer.py:5-56 summary of one trace -> float64[6] in ANALYTICS_KEYS order."""
    w = _f64(wealth)
    iv = np.ascontiguousarray(inventory, np.int32)
    tr = np.ascontiguousarray(np.asarray(is_trade).astype(bool), np.uint8)
    out = np.zeros(6, np.float64)
    scratch = np.empty(max(len(w), 1), np.float64)
    L = lib()
    L.oracle_analytics.restype = None
    L.oracle_analytics(C.c_int64(len(w)), _p(w, _f64p), _p(iv, _i32p), _p(tr, C.POINTER(C.c_uint8)),
                       _p(scratch, _f64p), _p(out, _f64p))
    return out
