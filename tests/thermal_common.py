"""Inputs shared by the non-isothermal tests: a synthetic 61-row CIE cooling table in the format of
tables/corocool.tab (which the reference does not ship, cooling.f90:71) and the problem set-up."""
import numpy as np


def cooling_table():
    """61 rows (log10 T, log10 Lambda): T from 10 K to 10^7 K in steps of 0.1 dex, a smooth curve with the
    steep rise of the Lyman-alpha cooling near 10^4 K and a slow rise above (shape only; not physical data)"""
    lt = 1.0 + 0.1 * np.arange(61)
    lc = -26.0 + 1.6 * np.tanh((lt - 4.15) * 3.0) + 0.25 * (lt - 4.0)
    return lt, lc


def setup_thermal_oracle(p, tables4, zred=9.0, cosmological=True, T0=None):
    from problems import setup_oracle
    o = setup_oracle(p, tables=tables4[:2])
    if T0 is not None:
        o.set_temperature(T0)
    o.set_isothermal(False)
    o.set_heat_tables(*tables4[2:])
    o.set_cooling_table(*cooling_table())
    o.set_redshift(zred, cosmological)
    return o


def setup_thermal_gpu(p, tables4, zred=9.0, cosmological=True, T0=None, **kw):
    from problems import setup_gpu
    e = setup_gpu(p, tables=tables4[:2], isothermal=0, cosmological=int(bool(cosmological)), **kw)
    if T0 is not None:
        e.set_temperature(T0)
    e.set_heat_tables(*tables4[2:])
    e.set_cooling_table(*cooling_table())
    e.set_redshift(zred)
    return e
