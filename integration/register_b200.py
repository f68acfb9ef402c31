#!/usr/bin/env python3
"""Register the B200 backend in a CUNQA checkout.

    python integration/register_b200.py --cunqa /path/to/cunqa [--dry-run]

Mechanical, idempotent edits -- nothing of the reference is stored in this repo, the script only inserts lines:
  * copies cunqa_b200/plugin/ (+ include/cunqa_b200.h) to src/backends/simulators/B200/ and adds the subdirectory;
  * utils/constants.hpp.in: SIMULATORS enum + SIMULATORS_MAP entry, SUPPORTED_{SIMPLE,CC,QC}_SIMULATORS, B200_BASIS_GATES;
  * utils/helpers/basis_gates.hpp: the B200 case;
  * cli/setup_qpus.cpp: includes + one case per backend family; cli/setup_executor.cpp: include + case;
  * cli/CMakeLists.txt: link the plugin; cli/qraise/*_conf_qraise.hpp: --gpu allow-list and a GPU_ARCH == 100 gres line.
After it: build libcunqa_b200.so (python -m cunqa_b200._build) and configure CUNQA with
-DCUNQA_B200_LIBRARY=<path>/libcunqa_b200.so -DCUNQA_B200_INCLUDE_DIR=<path>/include.
"""
import argparse
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASIS_GATES = ("measure reset id x y z h s sdg t tdg sx sxdg sy sydg u1 p gp rx ry rz u2 r u3 u swap iswap dcx ecr cx cy cz ch cs csdg ct "
               "csx csxdg cp cu1 crx cry crz rxx ryy rzz rzx xxmyy xxpyy cu2 cr cu3 cu ccx ccy ccz cswap cecr mcx mcy mcz mch mcs mct "
               "mcsx mcp mcu1 mcrx mcry mcrz mcu2 mcr mcu3 mcu mcswap unitary cunitary diagonal").split()


class Edit:
    def __init__(self, base, dry):
        self.base, self.dry, self.log = base, dry, []

    def path(self, rel):
        return os.path.join(self.base, rel)

    def apply(self, rel, fn, what):
        p = self.path(rel)
        if not os.path.exists(p):
            self.log.append(f"MISSING  {rel}: {what}")
            return
        old = open(p).read()
        new = fn(old)
        if new == old:
            self.log.append(f"ok       {rel}: {what} (already there)")
            return
        if not self.dry:
            open(p, "w").write(new)
        self.log.append(f"edited   {rel}: {what}")


def insert_before(text, anchor_regex, line, guard):
    """insert `line` before the first match of anchor_regex unless `guard` already occurs in text"""
    if guard in text:
        return text
    m = re.search(anchor_regex, text, flags=re.M)
    if not m:
        raise RuntimeError(f"anchor {anchor_regex!r} not found")
    return text[:m.start()] + line + text[m.start():]


def insert_after(text, anchor_regex, line, guard):
    if guard in text:
        return text
    m = re.search(anchor_regex, text, flags=re.M)
    if not m:
        raise RuntimeError(f"anchor {anchor_regex!r} not found")
    return text[:m.end()] + line + text[m.end():]


def constants(text):
    if not re.search(r"^\s*B200\s*$", text, flags=re.M):
        text = re.sub(r"^(\s*)CUNQASIM(\s*)$", r"\1CUNQASIM,\n\1B200\2", text, count=1, flags=re.M)
    text = insert_after(text, r'^\s*\{"Cunqa", CUNQASIM\},\s*$', '\n  {"B200", B200},', '{"B200", B200}')
    for fam in ("SIMPLE", "CC", "QC"):
        pat = rf'(SUPPORTED_{fam}_SIMULATORS = \{{[^}}]*?)"Cunqa"\}}'
        if not re.search(rf'SUPPORTED_{fam}_SIMULATORS = \{{[^}}]*"B200"', text):
            text = re.sub(pat, r'\1"Cunqa", "B200"}', text, count=1)
    gates = ", ".join(f'"{g}"' for g in BASIS_GATES)
    block = f"const std::vector<std::string> B200_BASIS_GATES = {{\n    {gates}\n}};\n\n"
    return insert_before(text, r"^const std::vector<std::string> AER_BASIS_GATES", block, "B200_BASIS_GATES")


def basis_gates(text):
    return insert_before(text, r"^\s*default:\s*$", "  case B200:\n      return B200_BASIS_GATES;\n      break;\n", "B200_BASIS_GATES")


def setup_qpus(text):
    inc = '#include "backends/simulators/B200/b200_plugin.hpp"\n'
    text = insert_before(text, r'^#include "utils/constants.hpp"', inc, "b200_plugin.hpp")
    fams = (("do not support simple simulation", "B200SimpleSimulator, SimpleConfig, SimpleBackend", "no_comm"),
            ("classical communication simulation", "B200CCSimulator, CCConfig, CCBackend", "classical_comm"),
            ("quantum communication simulation", "B200QCSimulator, QCConfig, QCBackend", "quantum_comm"))
    for needle, targs, comm in fams:
        cls = targs.split(",")[0]
        if f"turn_ON_QPU<{cls}" in text:
            continue
        m = re.search(rf"^(\s*)default:\s*\n\s*LOGGER_ERROR\(\"[^\"]*{re.escape(needle)}", text, flags=re.M)
        if not m:
            raise RuntimeError(f"switch for '{needle}' not found in setup_qpus.cpp")
        ind = m.group(1)
        case = (f'{ind}case murmur::hash("B200"):\n{ind}    LOGGER_DEBUG("QPU going to turn on with {cls}.");\n'
                f'{ind}    turn_ON_QPU<{targs}>(backend_json, mode, name, family, "{comm}");\n{ind}    break;\n')
        text = text[:m.start()] + case + text[m.start():]
    return text


def setup_executor(text):
    text = insert_before(text, r'^#include "utils/constants.hpp"|^#include "logger.hpp"', '#include "backends/simulators/B200/b200_plugin.hpp"\n', "b200_plugin.hpp")
    if "B200Executor" in text:
        return text
    m = re.search(r"^(\s*)default:\s*\n\s*LOGGER_ERROR\(\"Not a supported simulator", text, flags=re.M)
    if not m:
        raise RuntimeError("default case not found in setup_executor.cpp")
    ind = m.group(1)
    case = (f'{ind}case murmur::hash("B200"):\n{ind}{{\n{ind}    LOGGER_DEBUG("Raising executor with B200.");\n'
            f'{ind}    B200Executor executor(n_qpus);\n{ind}    executor.run();\n{ind}    break;\n{ind}}}\n')
    return text[:m.start()] + case + text[m.start():]


def cli_cmake(text):
    if "b200-simulator" in text:
        return text
    text = re.sub(r"(quest_simple_simulator quest_cc_simulator quest_qc_simulator)\)", r"\1\n                                         b200-simulator)", text, count=1)
    return re.sub(r"(quest_executor) (logger_executor\))", r"\1 b200-simulator \2", text, count=1)


def qraise(text):
    text = text.replace('simulators_with_gpu_support = {"Aer"};', 'simulators_with_gpu_support = {"Aer", "B200"};')
    if "GPU_ARCH == 100" in text:
        return text
    line = '#elif GPU_ARCH == 100\n    sbatchFile << "#SBATCH --gres=gpu:b200:" << std::to_string(args.n_qpus) << "\\n";\n'
    return re.sub(r"^#endif // GPU_ARCH", lambda m: line + m.group(0), text, flags=re.M)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--cunqa", required=True, help="root of the CUNQA checkout")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args(argv)
    e = Edit(args.cunqa, args.dry_run)
    dst = e.path("src/backends/simulators/B200")
    if not args.dry_run:
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(ROOT, "cunqa_b200", "plugin"))):
            shutil.copy(os.path.join(ROOT, "cunqa_b200", "plugin", f), os.path.join(dst, f))
        shutil.copy(os.path.join(ROOT, "include", "cunqa_b200.h"), os.path.join(dst, "cunqa_b200.h"))
    e.log.append("copied   src/backends/simulators/B200/ (plugin classes + cunqa_b200.h)")
    e.apply("src/backends/simulators/CMakeLists.txt", lambda t: t if "add_subdirectory(B200)" in t else t.rstrip("\n") + "\nadd_subdirectory(B200)\n", "add_subdirectory(B200)")
    e.apply("src/utils/constants.hpp.in", constants, "SIMULATORS / SIMULATORS_MAP / SUPPORTED_* / B200_BASIS_GATES")
    e.apply("src/utils/helpers/basis_gates.hpp", basis_gates, "case B200")
    e.apply("src/cli/setup_qpus.cpp", setup_qpus, "includes + simple / cc / qc cases")
    e.apply("src/cli/setup_executor.cpp", setup_executor, "include + executor case")
    e.apply("src/cli/CMakeLists.txt", cli_cmake, "link b200-simulator")
    for f in ("simple_conf_qraise.hpp", "cc_conf_qraise.hpp", "qc_conf_qraise.hpp"):
        e.apply(os.path.join("src/cli/qraise", f), qraise, "--gpu allow-list + GPU_ARCH == 100 gres")
    print("\n".join(e.log))
    return 0


if __name__ == "__main__":
    sys.exit(main())
