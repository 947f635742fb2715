"""FTPEnv -- the scalar First-Passage-Time env with the reference's exact surface.

Mirrors /root/reference/Env/market_env.py:3-67: ``FTPEnv(phi, tick_size, fee_rate)``,
``reset() -> (inventory, cash)``, ``step(action, mid_next, best_ask, best_bid, buy_max, sell_min,
adv_action=None) -> (reward, info)`` and the read/write attributes ``inventory cash i_max i_min phi
tick_size fee_rate``.  The arithmetic is the host instantiation of csrc/sgmm_step_core.h -- the
same header the CUDA kernels compile -- reached through ``sgmm_env_step_host`` of the C ABI.  This
class serves the per-bar Python loops of the blind test / backtest callers; population work goes
through :mod:`engine`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _attr(name, cast):
    return property(lambda self: cast(getattr(self._s, name)),
                    lambda self, v: setattr(self._s, name, cast(v)))


class FTPEnv:
    def __init__(self, phi=0.01, tick_size=0.01, fee_rate=0.0000):
        self._s = _lib.EnvState()
        self._info = _lib.StepInfo()
        self._act = (C.c_int64 * 2)()
        self._adv = (C.c_int64 * 2)()
        self._actf = (C.c_double * 2)()
        self._step = _lib.lib().sgmm_env_step_host
        self._step_real = _lib.lib().sgmm_env_step_host_real
        _lib.check(_lib.lib().sgmm_env_init(C.byref(self._s), float(phi), float(tick_size), float(fee_rate)))

    phi = _attr("phi", float)
    tick_size = _attr("tick_size", float)
    fee_rate = _attr("fee_rate", float)
    inventory = _attr("inventory", int)
    cash = _attr("cash", float)
    i_max = _attr("i_max", int)
    i_min = _attr("i_min", int)

    def reset(self):
        self._s.inventory = 0
        self._s.cash = 0.0
        return self.inventory, self.cash

    def step(self, action, mid_next, best_ask, best_bid, buy_max, sell_min, adv_action=None):
        a0, a1 = action[0], action[1]                             # market_env.py:23: offsets as given, floats included
        adv = None
        if adv_action is not None:
            d = np.round(adv_action).astype(int)                  # market_env.py:26
            self._adv[0], self._adv[1] = int(d[0]), int(d[1])
            adv = self._adv
        if float(a0).is_integer() and float(a1).is_integer() and abs(a0) < 2 ** 62 and abs(a1) < 2 ** 62:
            self._act[0], self._act[1] = int(a0), int(a1)
            _lib.check(self._step(C.byref(self._s), self._act, adv, float(mid_next), float(best_ask),
                                  float(best_bid), float(buy_max), float(sell_min), C.byref(self._info)))
        else:                                                     # e.g. an unrounded GLFT offset: quote at 1.7 ticks, not 1
            self._actf[0], self._actf[1] = float(a0), float(a1)
            _lib.check(self._step_real(C.byref(self._s), self._actf, adv, float(mid_next), float(best_ask),
                                       float(best_bid), float(buy_max), float(sell_min), C.byref(self._info)))
        i = self._info
        info = {'pnl_reward': i.pnl_reward, 'inventory_reward': i.inventory_reward,
                'fee_paid': i.fee_paid, 'fill_buy': i.fill_buy, 'fill_sell': i.fill_sell}
        return np.float64(i.reward), info
