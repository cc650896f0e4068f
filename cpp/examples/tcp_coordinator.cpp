// tcp_coordinator: see pipeline_cli.hpp (reference: examples/tcp_coordinator.cpp).
#include "pipeline_cli.hpp"

int main(int argc, char** argv) { return tnn::cli::coordinator_main(argc, argv, false); }
