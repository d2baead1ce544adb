"""clearsky_b200 -- host-side mirror of ClearSky.jl's API for the line-by-line radiative-transfer hot path,
backed by libclearsky_b200.so (hand-written sm_100a CUDA, FP64).  The Julia wrapper
(../julia/ClearSkyB200.jl) issues the same C-ABI calls via ccall; this Python twin exists because no Julia
runtime is available in the build/test image.  There is no CPU fallback anywhere in this package.
"""
from . import constants
from ._lib import ClearSkyError, Context, default_context, device_count, LIB_PATH
from .absorbers import (AcceleratedAbsorber, SigmaWorkspace, UnifiedAbsorber, getwavenumbers, pressurelimits,
                        temperaturelimits, unifyabsorbers)
from .atmospherics import DryAdiabat
from .cia import CIA, CIATables, readcia
from .core import B200Discretized, Discretized, FluxPack
from .fluxes import (fluxes, fluxes_batch, monochromaticfluxes, monochromaticfluxes_, netfluxes, opticaldepth, radiate, radiate_,
                     transmittance)
from .gases import AtmosphericDomain, Gas, GrayGas, LineGas, SemiGrayGas
from .line_shapes import (PHCO2, PHCO2_b200_inplace, DeviceLines, device_lines, doppler, doppler_b200_inplace,
                          lorentz, lorentz_b200_inplace, scaleintensity, voigt, voigt_b200_inplace, xsec, αdoppler,
                          γlorentz)
from .molparam import MOLPARAM, TMAX, TMIN
from .par import SpectralLines, parse_records_b200, readpar, readpar_b200, writepar
from .quadrature import lobattonodes, streamnodes
from .radau import Radau, outgoing
from .rcm import RCM
from .sharding import DeviceGroup, ShardedAbsorber, ShardedLineByLine, sharded_fluxes
from .util import AtmosphericProfile, chebygrid, pressuregrid, trapz
