"""Debug helper: runs the FFI lifecycle on the first reference fixture in-process with tracing."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler; faulthandler.enable()
from oracle import plonk as pl
from tests.ffi_util import write_srs_file
from tests.test_plonk_oracle import FIXTURES
home = tempfile.mkdtemp()
write_srs_file(home, pl.SRS(128, 0xB2000005))
os.environ["XDG_CONFIG_HOME"] = home
os.environ["B200ZK_FFI_TRACE"] = "1"
from noir_backend_using_gnark_b200 import ffi
js, vals = FIXTURES[0]
pk, vk = ffi.preprocess(js, 5)
print(len(pk), len(vk))
