"""Values of the reference's Fortran parameters as gfortran stores them.

Most "double" constants of C2-Ray3Dm are default-real (binary32) literals widened to real(dp)
(SURVEY Appendix B); `_f` reproduces that.  Citations are reference file:line.
"""
import numpy as np


def _f(x):
    """default-real literal widened to double"""
    return float(np.float32(x))


pi = _f(3.141592654)                                   # mathconstants.f90:21
m_p = 1.672661e-24                                     # cgsconstants.f90:26
c = 2.997925e+10                                       # :28
hplanck = 6.6260755e-27                                # :30
sigma_SB = 5.670e-5                                    # :32
k_B = 1.381e-16                                        # :34
G_grav = 6.6732e-8                                     # :36
ev2k = float(np.float32(1.0) / np.float32(8.617e-05))  # :39
ev2fr = _f(0.241838e15)                                # :53
two_pi_over_c_square = _f(2.0) * pi / (c * c)          # :61
albpow = -0.7                                          # :64
bh00 = 2.59e-13                                        # :66
eth0 = _f(13.598)                                      # :76
temph0 = eth0 * ev2k                                   # :80
colh0 = _f(1.3e-8) * _f(0.83) * _f(1.0) / (eth0 * eth0)  # :86
ethe = (_f(24.587), _f(54.416))                        # :101
sigma_HI_at_ion_freq = 1.0 * _f(6.30e-18)              # cgsphotoconstants.f90:24
ion_freq_HI = ev2fr * eth0                             # :31
ion_freq_HeII = ev2fr * ethe[1]                        # :33
abu_he = _f(0.074)                                     # abundances.f90:23
abu_c = _f(7.1e-7)                                     # :26
mu = (_f(1.0) - abu_he) + _f(4.0) * abu_he             # :32
R_SOLAR = _f(6.9599e10)                                # cgsastroconstants.f90:23
YEAR = _f(3.15576E+07)                                 # :27
pc = _f(3.086e18)                                      # :29
Mpc = _f(1e6) * pc                                     # :31
h = _f(0.7)                                            # cosmoparms.f90:28
Omega0 = _f(0.27)                                      # :30
Omega_B = _f(0.044)                                    # :31
H0 = h * _f(100.0) * _f(1e5) / Mpc                     # :41
rho_crit_0 = _f(3.0) * H0 * H0 / (_f(8.0) * pi * G_grav)  # :42
bb_Teff = _f(5.0e4)                                    # sed_parameters.f90:29
bb_S_star = 1e48                                       # :31
bb_MinFreq = ion_freq_HI                               # :35
bb_MaxFreq = ion_freq_HeII * 10.00                     # :36
pl_index_cross_section_HI = 2.8                        # radiation_sizes.f90:85
xh_initial = _f(2e-4)                                  # ionfractions_module.F90:49
initial_temperature = _f(1e4)                          # c2ray_parameters.f90:116
boxsize_test = _f(100.0)                               # nbody_test.F90:44 (Mpc/h)
