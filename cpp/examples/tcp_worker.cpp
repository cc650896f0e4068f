// tcp_worker: a pipeline / data-parallel stage reachable over the TCP bus (reference: examples/tcp_worker.cpp:39-141).
//   tcp_worker <port> [--gpu [index]] [--host 127.0.0.1] [--io-threads N] [--num-threads N]
#include <cstring>
#include <iostream>

#include "tnn/distributed.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  int port = 0, gpu = -1;
  std::string host = "127.0.0.1";
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--gpu")) gpu = (i + 1 < argc && isdigit((unsigned char)argv[i + 1][0])) ? atoi(argv[++i]) : 0;
    else if (!strcmp(argv[i], "--host") && i + 1 < argc) host = argv[++i];
    else if ((!strcmp(argv[i], "--io-threads") || !strcmp(argv[i], "--num-threads")) && i + 1 < argc) ++i;
    else if (isdigit((unsigned char)argv[i][0])) port = atoi(argv[i]);
    else if (!strcmp(argv[i], "--help")) {
      std::cout << "usage: tcp_worker <port> [--gpu [index]] [--host ADDR]\n";
      return 0;
    }
  }
  try {
    const Device dev = gpu >= 0 && DeviceManager::getInstance().has_gpu() ? getGPU(gpu) : getHost();
    auto comm = std::make_unique<TCPCommunicator>(host, port, 4, &dev, "tcp-worker-" + std::to_string(port));
    Worker w(std::move(comm), dev, "tcp-worker-" + std::to_string(port));
    std::cout << "[tcp_worker] listening on " << host << ":" << port << " device " << dev.name() << std::endl;
    w.start(false);       // blocks until SHUTDOWN
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[tcp_worker] error: " << e.what() << std::endl;
    return 1;
  }
}
