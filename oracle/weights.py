"""TEST INFRASTRUCTURE — shim: the seeded weight / input generators live in vcg_b200/synthetic.py (they generate data, they
do not restate the algorithm, and bench.py's measured arm needs them too); re-exported here under their historical name
for the oracle scripts and the tests.  Loaded by file path, not through the package: the golden scripts must keep this
repo's `model` / `data` packages OFF sys.path so that the names resolve to the unmodified reference's."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video-chapter-generation_b200",
                     "vcg_b200", "synthetic.py")
_spec = importlib.util.spec_from_file_location("vcg_b200_synthetic_standalone", _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
