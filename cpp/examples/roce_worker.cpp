// roce_worker: see pipeline_cli.hpp (reference: examples/roce_worker.cpp).
#include "pipeline_cli.hpp"

int main(int argc, char** argv) { return tnn::cli::worker_main(argc, argv, true); }
