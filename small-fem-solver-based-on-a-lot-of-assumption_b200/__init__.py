"""jacket_b200 -- B200-native (sm_100a) implementation of the jacket tool's hot path:
Morison load integration -> Timoshenko assembly -> factor-once multi-RHS FP64 solve ->
reactions / member forces / utilisation -> critical-phase reduction.

The class surface mirrors the reference's analysis module
(JacketAnalysisGUI_v2.py:115-803); the arithmetic runs in hand-written CUDA kernels
(csrc/) behind the C ABI of include/jacket_b200.h.  There is no CPU fallback.
"""
from .sections import TubularSection
from .structure import CustomJacketStructure, create_default_3leg_jacket, generate_jacket, default_sections
from .wave import RaschiiWave, g, enable_nonlinear_waves
from . import wavefit
from .morison import MorisonCalculator, phase_times
from .fem import FEMSolver, BeamElement3D
from .analysis import (AnalysisParams, PhaseScanResult, EnsembleResult, ensemble_scan, dispersion_wavenumbers, phase_scan, phase_scan_from_params, run_analysis,
                       static_load, interface_load_vector, apply_self_weight, build_structure)
from .engine import Engine, get_engine
from ._lib import JacketError, NotPositiveDefinite, TABLE_COLUMNS, MEMBER_COLUMNS, DETAIL_COLUMNS, LIB_PATH

DEFAULT_RHO_WATER = 1025
DEFAULT_E = 210000
DEFAULT_NU = 0.3
DEFAULT_FY = 355
DEFAULT_RHO_STEEL = 7850

__all__ = [n for n in dir() if not n.startswith("_")]
