#!/usr/bin/env bash
# Sanitizer build of the C++ host (SAN=thread, the default, or SAN=address,undefined): (cpp/src + cpp/tests, linked with the un-instrumented sm_100a kernel objects) and a run of
# the multi-threaded distributed tests on the CPU: in-process / TCP / rendezvous pipelines, the host-slab rendezvous, fault handling.
#   tools/sanitize_cpp_host.sh [test-name-substring ...]        (default: pipeline roce lost_microbatch utils_threading)
set -euo pipefail
cd "$(dirname "$0")/.."
san="${SAN:-thread}"
out="/tmp/tnn_san_host_${san//,/_}"
mkdir -p "$out"
export ASAN_OPTIONS="${ASAN_OPTIONS:-detect_leaks=0:protect_shadow_gap=0}" UBSAN_OPTIONS="${UBSAN_OPTIONS:-print_stacktrace=1}"
cuda=${CUDA_HOME:-/usr/local/cuda}
flags="-O1 -g -std=c++17 -fsanitize=$san -fno-omit-frame-pointer -fPIC -Icpp/include -Itnn_b200/csrc -I$cuda/include -pthread -Wno-unused-function"
pids=()
for s in cpp/src/*.cpp cpp/tests/tnn_tests.cpp; do
  o="$out/$(basename "${s%.cpp}").o"
  # rebuild when the source or ANY header is newer than the object
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find cpp/include tnn_b200/csrc \( -name '*.hpp' -o -name '*.cuh' \) -newer "$o" -print -quit)" ]; then
    g++ $flags -c "$s" -o "$o" & pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
g++ -fsanitize=$san -o "$out/tnn_tests_san" "$out"/*.o tnn_b200/_build/*.cu.o -L"$cuda/lib64" -lcudart -Wl,-rpath,"$cuda/lib64" -pthread -ldl
rc=0
for t in "${@:-pipeline roce lost_microbatch utils_threading}"; do
  for name in $t; do
    echo "=== $san sanitizer: tnn_tests $name"
    TSAN_OPTIONS="halt_on_error=0 second_deadlock_stack=1" "$out/tnn_tests_san" "$name" 2>&1 | grep -v "^\[ SKIPPED" | tail -40 || rc=1
  done
done
exit $rc
