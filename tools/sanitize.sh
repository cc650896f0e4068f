#!/bin/bash
# Sanitizer targets (the reference only sets ASan/UBSan flags in CPU debug builds and
# -fdevice-sanitize=memcheck in CUDA debug flags; it has no compute-sanitizer or TSan target).
#
#   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [pytest -k expr]   # on a GPU box
#   tools/sanitize.sh tsan                                                       # host runtime (TCP bus) under TSan
set -euo pipefail
cd "$(dirname "$0")/.."
mode="${1:-memcheck}"; shift || true
case "$mode" in
  memcheck|racecheck|synccheck|initcheck)
    expr="${1:-linear_fwd or conv or batchnorm or attention}"
    exec compute-sanitizer --tool "$mode" --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_kernels_gpu.py -x -q -k "$expr"
    ;;
  tsan)
    # stand-alone TSan build of the message bus + allocators (no Python: libtsan cannot be preloaded late)
    out=/tmp/tnn_tsan_test
    cat > /tmp/tnn_tsan_main.cpp <<'CPP'
#include "runtime/tcp_bus.hpp"
#include "runtime/allocator.hpp"
#include <cassert>
#include <cstdio>
int main() {
  tnn::TcpBus a(0, 4, 1 << 16), b(0, 4, 1 << 16);
  a.start(); b.start();
  a.connect("127.0.0.1", b.port());
  std::vector<uint8_t> big(3 << 20, 7);
  big[0] = 5; big[1] = 0;
  std::thread t1([&] { for (int i = 0; i < 20; ++i) a.send("127.0.0.1:" + std::to_string(b.port()), big.data(), big.size()); });
  std::thread t2([&] { for (int i = 0; i < 20; ++i) a.send("127.0.0.1:" + std::to_string(b.port()), big.data(), 64); });
  int got = 0; std::string from; std::vector<uint8_t> msg;
  while (got < 40 && b.recv(5000, &from, &msg)) ++got;
  t1.join(); t2.join();
  a.stop(); b.stop();
  std::printf("received %d/40 messages\n", got);
  return got == 40 ? 0 : 1;
}
CPP
    g++ -std=c++17 -O1 -g -fsanitize=thread -Itnn_b200/csrc /tmp/tnn_tsan_main.cpp -o "$out" -pthread
    exec "$out"
    ;;
  *) echo "unknown mode $mode"; exit 2;;
esac
