"""
Physical constants and host-side scalar coefficients of the LOS integrator (CGS).
The literals are frozen and identical to oracle/los.py and csrc/los.cu.
"""
import math

C_LIGHT = 2.99792458e10        # cm / s
Q_E = 4.80320471e-10           # esu
M_E = 9.10938356e-28           # g
K_B = 1.38064852e-16           # erg / K
GEV_ERG = 1.6021766208e-3      # erg
KPC_CM = 3.0856775814913673e21  # cm
MEC2_ERG = M_E * C_LIGHT * C_LIGHT
RM_CONST = 0.812               # rad m^-2 / (cm^-3 uG pc), tutorials/tutorial_simulator.ipynb:81


def sync_coefficient(p, nu_ghz, e0_gev):
    """
    K(p, nu): Rayleigh-Jeans brightness temperature [K] per unit
    (cre [GeV m^2 s sr]^-1) * (Bperp[uG]^2)^((p+1)/4) * cm for a power-law CRE spectrum of
    index p normalised at e0_gev (Rybicki & Lightman eq. 6.36).
    """
    nu = nu_ghz * 1.0e9
    t_conv = C_LIGHT * C_LIGHT / (2.0 * K_B * nu * nu)
    ptot = (math.sqrt(3.0) * Q_E ** 3 / (2.0 * math.pi * MEC2_ERG * (p + 1.0))
            * math.gamma(p / 4.0 + 19.0 / 12.0) * math.gamma(p / 4.0 - 1.0 / 12.0)
            * (2.0 * math.pi * M_E * C_LIGHT * nu / (3.0 * Q_E)) ** (-(p - 1.0) / 2.0))
    c_gamma = ((4.0 * math.pi / C_LIGHT) * (e0_gev * GEV_ERG) ** p
               * MEC2_ERG ** (1.0 - p) / (GEV_ERG * 1.0e4))
    return t_conv * 0.5 * ptot * c_gamma * (1.0e-6) ** ((p + 1.0) / 2.0)


H_PLANCK = 6.62607015e-27      # erg s
T_CMB = 2.725                  # K


def rj_to_cmb(nu_ghz, t_cmb=T_CMB):
    """
    Factor from Rayleigh-Jeans brightness temperature (what the kernels integrate) to thermodynamic CMB
    temperature: dT_CMB = dT_RJ (e^x - 1)^2 / (x^2 e^x), x = h nu / (k T_cmb).  The reference's key
    convention is K-cmb (observables/observable_dict.py:15-19) and the hammurabiX output recorded in
    tutorials/tutorial_hammurabi.ipynb is reproduced to 3e-5 with this factor (1.01375 at 23 GHz) --
    tools/hammurabi_tutorial_pin.py.
    """
    x = H_PLANCK * nu_ghz * 1.0e9 / (K_B * t_cmb)
    ex = math.exp(x)
    return (ex - 1.0) ** 2 / (x * x * ex)
