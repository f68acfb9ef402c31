#!/usr/bin/env bash
# Builds the reference's own boundary code into oracle/_ref/ (git-ignored, travels with gpurun):
#   1. oracle/_ref/ref_parse   = /root/reference/src/{quantum_task.cpp,utils/json.cpp} + ref_parse.cpp
#   2. compile check of the plugin classes (cunqa_b200/plugin/*.cpp) against the reference's headers
#   3. oracle/_ref/plugin_run = the plugin classes LINKED with the reference's quantum_task.cpp + utils/json.cpp,
#      an in-process ClassicalChannel (fake_channel.cpp) and libcunqa_b200.so: runs execute() end to end
# Only gcc on the reference sources where they lie; the reference's CMake build is NOT run (it fetches
# Aer, ZMQ, ... over the network).  No reference source is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF=/root/reference/src
OUT="$ROOT/oracle/_ref"
SITE="$(python -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')"
NLOHMANN="$SITE/include/cudnn_frontend/thirdparty"
SPDLOG="$SITE/flashinfer/data/spdlog/include"
[ -d "$REF" ] || { echo "no /root/reference: keeping the prebuilt oracle/_ref"; exit 0; }
[ -f "$NLOHMANN/nlohmann/json.hpp" ] && [ -d "$SPDLOG/spdlog" ] || { echo "nlohmann/spdlog headers not found: skipping"; exit 0; }
mkdir -p "$OUT/include/utils"
sed 's#@CMAKE_INSTALL_PREFIX@#/opt/cunqa#' "$REF/utils/constants.hpp.in" > "$OUT/include/utils/constants.hpp"
INC=(-I"$OUT/include" -I"$REF" -I"$REF/utils" -I"$REF/utils/logger" -I"$REF/backends" -I"$NLOHMANN" -I"$SPDLOG" -I"$ROOT/include" -I"$ROOT/cunqa_b200/plugin")
STAMP="$OUT/.stamp"
NEWEST=$(ls -t "$HERE"/*.cpp "$ROOT"/cunqa_b200/plugin/*.?pp "$ROOT"/include/*.h "$HERE"/build.sh | head -1)
if [ -x "$OUT/ref_parse" ] && [ -x "$OUT/plugin_run" ] && [ -f "$STAMP" ] && [ "$STAMP" -nt "$NEWEST" ]; then echo "oracle/_ref up to date"; exit 0; fi
g++ -std=c++20 -O1 -w "${INC[@]}" "$HERE/ref_parse.cpp" "$REF/quantum_task.cpp" "$REF/utils/json.cpp" -o "$OUT/ref_parse"
for f in "$ROOT"/cunqa_b200/plugin/*.cpp; do
  STORE=/tmp g++ -std=c++20 -w -fsyntax-only "${INC[@]}" "$f"
done
# 3. the plugin LINKED and runnable: product plugin + the reference's task parser / registry + an in-process
#    ClassicalChannel + libcunqa_b200.so (resolved relative to the binary, so it also runs from the gpurun snapshot)
if [ -f "$ROOT/cunqa_b200/libcunqa_b200.so" ]; then
  g++ -std=c++20 -O1 -w -pthread "${INC[@]}" -I"$REF/classical_channel" \
      "$HERE/plugin_run.cpp" "$HERE/fake_channel.cpp" "$ROOT"/cunqa_b200/plugin/b200_plugin.cpp \
      "$REF/quantum_task.cpp" "$REF/utils/json.cpp" \
      -L"$ROOT/cunqa_b200" -lcunqa_b200 -Wl,-rpath,'$ORIGIN/../../cunqa_b200' -o "$OUT/plugin_run"
fi
echo "plugin sources compile against the reference headers" > "$OUT/plugin_check.txt"
touch "$STAMP"
echo "built $OUT/ref_parse ; plugin compile check ok"
