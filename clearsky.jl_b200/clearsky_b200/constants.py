"""Physical constants, bit-exact copies of the reference's values.

Reference: src/constants.jl:1-27 (note k is the 2014 CODATA value on purpose),
src/absorption/line_shapes.jl:2-5, src/hitran/molparam.jl:1-2.
"""
import math

c = 299792458.0        # constants.jl:2   speed of light [m/s]
h = 6.62607015e-34     # constants.jl:4   Planck [J s]
k = 1.38064852e-23     # constants.jl:6   Boltzmann [J/K]
sigma_sb = 5.67037442e-8   # constants.jl:8
R = 8.31446262         # constants.jl:10  gas constant [J/K/mole]
A = 101325.0           # constants.jl:12  Pa per atm
Na = 6.02214076e23     # constants.jl:14
Lo2 = 7.21879268e38    # constants.jl:20  Loschmidt^2 [molecules^2/cm^6]
Tref = 296.0           # constants.jl:23
T0 = 273.15            # constants.jl:25
Pmin = 1e-9            # constants.jl:27

TMIN = 25.0            # molparam.jl:1
TMAX = 1000.0          # molparam.jl:2

sqpi = math.sqrt(math.pi)                          # line_shapes.jl:2
osqpiln2 = 1 / math.sqrt(math.pi / math.log(2.0))  # line_shapes.jl:3
sqln2 = math.sqrt(math.log(2.0))                   # line_shapes.jl:4
c2 = 100.0 * h * c / k                             # line_shapes.jl:5
