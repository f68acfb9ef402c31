"""integration/register_b200.py applied to a scratch copy of the reference's registration sites.
Runs only where /root/reference is mounted (this container); nothing of the reference is stored in the repo."""
import os
import shutil
import subprocess
import sys
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SITES = ["src/backends/simulators/CMakeLists.txt", "src/utils/constants.hpp.in", "src/utils/helpers/basis_gates.hpp",
         "src/cli/setup_qpus.cpp", "src/cli/setup_executor.cpp", "src/cli/CMakeLists.txt",
         "src/cli/qraise/simple_conf_qraise.hpp", "src/cli/qraise/cc_conf_qraise.hpp", "src/cli/qraise/qc_conf_qraise.hpp"]

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not mounted")


def _apply(dst):
    return subprocess.run([sys.executable, os.path.join(ROOT, "integration", "register_b200.py"), "--cunqa", dst],
                          capture_output=True, text=True, check=True).stdout


def test_registration_edits_apply_once_and_compile(tmp_path):
    dst = str(tmp_path / "cunqa")
    for rel in SITES:
        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
        shutil.copy(os.path.join(REF, rel), os.path.join(dst, rel))
    first = _apply(dst)
    assert first.count("edited") == len(SITES) and "MISSING" not in first
    snapshot = {rel: open(os.path.join(dst, rel)).read() for rel in SITES}
    second = _apply(dst)
    assert "edited" not in second and second.count("already there") == len(SITES)
    assert snapshot == {rel: open(os.path.join(dst, rel)).read() for rel in SITES}  # idempotent
    qp = snapshot["src/cli/setup_qpus.cpp"]
    for cls in ("B200SimpleSimulator, SimpleConfig, SimpleBackend", "B200CCSimulator, CCConfig, CCBackend", "B200QCSimulator, QCConfig, QCBackend"):
        assert qp.count(f"turn_ON_QPU<{cls}>") == 1
    assert snapshot["src/cli/setup_executor.cpp"].count("B200Executor executor(n_qpus);") == 1
    assert os.path.exists(os.path.join(dst, "src/backends/simulators/B200/b200_plugin.cpp"))
    # the patched constants + basis-gate lookup compile and answer for "B200"
    site = sysconfig.get_paths()["purelib"]
    nl = os.path.join(site, "include", "cudnn_frontend", "thirdparty")
    if not os.path.exists(os.path.join(nl, "nlohmann", "json.hpp")):
        pytest.skip("nlohmann headers not found")
    inc = tmp_path / "inc" / "utils"
    os.makedirs(inc)
    (inc / "constants.hpp").write_text(snapshot["src/utils/constants.hpp.in"].replace("@CMAKE_INSTALL_PREFIX@", "/opt/cunqa"))
    shutil.copy(os.path.join(dst, "src/utils/helpers/basis_gates.hpp"), tmp_path / "basis_gates_patched.hpp")
    (tmp_path / "t.cpp").write_text('#include <cstdio>\n#include "basis_gates_patched.hpp"\n'
                                    'int main(){ auto g = get_basis_gates("B200"); std::printf("%zu %s\\n", g.size(), g[0].c_str());'
                                    ' return cunqa::constants::SIMULATORS_MAP.at("B200") == cunqa::constants::B200 ? 0 : 1; }\n')
    exe = str(tmp_path / "t")
    subprocess.run(["g++", "-std=c++20", "-w", f"-I{tmp_path / 'inc'}", f"-I{tmp_path / 'inc' / 'utils'}", f"-I{REF}/src", f"-I{REF}/src/utils",
                    f"-I{nl}", f"-I{tmp_path}", str(tmp_path / "t.cpp"), "-o", exe], check=True, env=dict(os.environ, STORE="/tmp"))
    out = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, STORE="/tmp"))
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import register_b200
    assert out.returncode == 0 and out.stdout.split() == [str(len(register_b200.BASIS_GATES)), "measure"], out
