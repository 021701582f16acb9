"""Import alias: the package directory is `scip-sdp_b200/` (not a valid Python identifier), so this stub package
extends its search path to that directory.  `import scip_sdp_b200.abi` loads `scip-sdp_b200/abi.py`."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "scip-sdp_b200"))
