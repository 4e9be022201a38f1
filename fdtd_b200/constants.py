"""Physical constants, same values and names as the reference (fdtd/constants.py:1-24)."""
from math import pi

c: float = 299792458.0
X: int = 0
Y: int = 1
Z: int = 2
mu0: float = 4e-7 * pi
eps0: float = 1.0 / (mu0 * c ** 2)
eta0: float = mu0 * c
