#!/usr/bin/env bash
# Build everything in-tree (reference: build.sh --clean / --debug / --verbose).  The kernels are sm_100a only; nvcc
# cross-compiles them without a GPU.
#   ./build.sh                 kernels + Python extension (tnn_b200/_C.so) + C++ host (cpp/lib/libtnn_b200.so, cpp/bin/*)
#   ./build.sh --clean         drop every built artefact first
#   ./build.sh --no-host       kernels + Python extension only
#   ./build.sh --cmake         configure and build the C++ host through cpp/CMakeLists.txt (into cpp/_cmake) instead
#   ./build.sh --debug         ASan + UBSan build of the C++ host and a run of its CPU tests (tools/sanitize_cpp_host.sh)
#   ./build.sh --tsan          ThreadSanitizer build of the C++ host and a run of its threaded tests
#   ./build.sh --test          after building: C++ unit tests and the CPU pytest suite
#   ./build.sh --verbose       print every compile / link step
set -euo pipefail
cd "$(dirname "$0")"
clean=0 host=1 cmake=0 debug=0 tsan=0 tests=0 verbose=False
while [[ $# -gt 0 ]]; do
  case "$1" in
    --clean | -c) clean=1 ;;
    --no-host) host=0 ;;
    --cmake) cmake=1 ;;
    --debug | -d) debug=1 ;;
    --tsan) tsan=1 ;;
    --test | -t) tests=1 ;;
    --verbose | -v) verbose=True ;;
    --help | -h) sed -n '2,11p' "$0" | sed 's/^# \{0,1\}//'; exit 0 ;;
    *) echo "unknown option: $1 (see --help)" >&2; exit 1 ;;
  esac
  shift
done
if [[ $clean == 1 ]]; then
  rm -rf ./tnn_b200/_build ./cpp/_build ./cpp/_cmake ./cpp/bin ./cpp/lib
  rm -f ./tnn_b200/_C.so ./tnn_b200/_host.so
  echo "[build.sh] cleaned"
fi
python - <<PY
from tnn_b200 import build as b
b.build(verbose=$verbose)
if $host and not $cmake:
    b.build_host(verbose=$verbose)
print("[build.sh] built:", b.ROOT / "_C.so")
PY
if [[ $cmake == 1 ]]; then
  cmake -S cpp -B cpp/_cmake -DCMAKE_BUILD_TYPE=Release
  cmake --build cpp/_cmake -j"$(nproc)"
fi
if [[ $debug == 1 ]]; then SAN=address,undefined tools/sanitize_cpp_host.sh; fi
if [[ $tsan == 1 ]]; then SAN=thread tools/sanitize_cpp_host.sh; fi
if [[ $tests == 1 ]]; then
  [[ $host == 1 && $cmake == 0 ]] && ./cpp/bin/tnn_tests | tail -3
  python -m pytest tests -q -m "not gpu"
fi
