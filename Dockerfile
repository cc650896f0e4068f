# Container image of the C++ host (reference: Dockerfile + docker-compose.yml -- CPU / multi-container pipeline runs).
# Build context = repository root.  The image needs the CUDA toolkit to compile the sm_100a kernels; the binaries run on
# the CPU when no GPU is visible (DEVICE_TYPE=CPU) and on B200 with `--gpus all`.
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04
RUN apt-get update && apt-get install -y --no-install-recommends cmake g++ make && rm -rf /var/lib/apt/lists/*
WORKDIR /app
COPY tnn_b200/csrc tnn_b200/csrc
COPY cpp cpp
RUN cmake -S cpp -B build -DCMAKE_BUILD_TYPE=Release && cmake --build build -j"$(nproc)" \
    && mkdir -p bin && cp build/trainer build/inferencer build/gpt2_inference build/tcp_worker build/tcp_coordinator \
       build/legacy_trainer build/graph_test build/msequential_example build/test_del build/device_manager_example \
       build/hardware_info_example build/bench_cpp build/tnn_tests bin/ && cp build/libtnn_b200.so /usr/local/lib/ && ldconfig
ENV DEVICE_TYPE=CPU MODEL_NAME=cifar10_resnet9 BATCH_SIZE=64 EPOCHS=1
CMD ["./bin/tnn_tests"]
