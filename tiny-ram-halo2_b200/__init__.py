"""tiny-ram-halo2_b200: B200 (sm_100a) backend for the data-parallel core of the halo2/IPA prover that proves the
TinyRAM circuit of Orbis-Tertius/tiny-ram-halo2 -- MSM, NTT / coset FFT and the EvaluationDomain transforms.

The product is the C-ABI shared library ``libtrp.so`` (include/tr_prover.h, sources in csrc/).  This Python
package is the host-side mirror of the halo2_proofs entry points the reference reaches from
/root/reference/src/test_utils.rs:21-49 (``best_multiexp``, ``best_fft``, ``EvaluationDomain``, ``Params.commit*``),
bound to the library with ctypes.  There is NO CPU fallback: importing works anywhere, but every operation raises
``TrpError`` unless libtrp.so is built and a CUDA device is present.

The directory name contains '-', so import it through ``__graft_entry__.load_package()`` (registers the package
as ``tiny_ram_halo2_b200``) or put the repo root on sys.path and call that helper.
"""
from ._lib import TrpError, Context, lib_path, load_library, PALLAS, VESTA, build_library  # noqa: F401
from .arithmetic import best_multiexp, best_fft, best_fft_group, hash_to_curve, Bases  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .commitment import Params  # noqa: F401
from .poly import Evaluator, new_evaluator  # noqa: F401
from . import permutation, lookup, ipa, plonk, verifier  # noqa: F401

__all__ = ["TrpError", "Context", "Bases", "best_multiexp", "best_fft", "best_fft_group", "hash_to_curve", "EvaluationDomain", "Params", "Evaluator", "new_evaluator",
           "PALLAS", "VESTA", "lib_path", "load_library", "build_library"]
