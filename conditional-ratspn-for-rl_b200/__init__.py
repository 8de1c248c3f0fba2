"""B200-native RAT-SPN / CSPN circuit evaluation behind the reference's module API.

The directory name contains hyphens, so import it through the alias module at the repository root:

    import ratspn_b200                     # -> this package
    from ratspn_b200 import RatSpn, RatSpnConfig, CSPN, CspnConfig, RatNormal

``install_as_reference_modules()`` additionally registers the sub-modules under the reference's flat module names
(``rat_spn``, ``cspn``, ``layers``, ``distributions``, ``utils``, ``type_checks``) so that unmodified caller code
(`from rat_spn import RatSpn`) picks up this implementation.
"""
import sys

from . import _abi, ops  # noqa: F401
from . import type_checks, utils, layers, distributions, rat_spn, cspn, diffsample, parallel, graphs  # noqa: F401
from . import sb3_adapter  # noqa: F401
from .graphs import GraphedCall  # noqa: F401
from .cspn import CSPN, CspnConfig, print_cspn_params  # noqa: F401
from .distributions import IndependentMultivariate, Leaf, RatNormal, truncated_normal_  # noqa: F401
from .layers import CrossProduct, Product, Sum  # noqa: F401
from .rat_spn import RatSpn, RatSpnConfig  # noqa: F401
from .utils import Sample, provide_evidence  # noqa: F401

__version__ = "0.1.0"


def install_as_reference_modules():
    """Make ``import rat_spn`` / ``cspn`` / ``layers`` / ``distributions`` / ``utils`` / ``type_checks`` resolve to
    this package (drop-in for code written against the reference's flat layout)."""
    for name in ("rat_spn", "cspn", "layers", "distributions", "utils", "type_checks"):
        sys.modules[name] = sys.modules[f"{__name__}.{name}"]
