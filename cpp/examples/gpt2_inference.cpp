// gpt2_inference: greedy text generation with the GPT-2 zoo models (reference: examples/gpt2_inference.cpp:19-127).
//   gpt2_inference [--model gpt2_small] [--model-path ckpt] [--vocab vocab.bin] [--prompt "text" | --ids 1,2,3]
//                  [--max-new 50] [--batch 1] [--device GPU:0] [--precision bf16|fp8] [--graph] [--bench B L steps]
// Unlike the reference (whole-model recompute per token, host arg-max over 103 MB of logits) generation runs over a KV
// cache with the arg-max on the device.  Without a vocabulary file prompts are given as ids and ids are printed.
#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>

#include "runtime/tokenizer.hpp"
#include "tnn/example_models.hpp"
#include "tnn/inference.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  try {
    std::string model = "gpt2_small", model_path, vocab, prompt, ids_arg, device = "", precision = "bf16";
    size_t max_new = 50, batch = 1, bench_b = 0, bench_l = 0, bench_steps = 0;
    bool graph = false;
    for (int i = 1; i < argc; ++i) {
      auto next = [&](const char* flag) -> std::string {
        if (i + 1 >= argc) throw std::runtime_error(std::string("missing value for ") + flag);
        return argv[++i];
      };
      if (!strcmp(argv[i], "--model")) model = next("--model");
      else if (!strcmp(argv[i], "--model-path")) model_path = next("--model-path");
      else if (!strcmp(argv[i], "--vocab")) vocab = next("--vocab");
      else if (!strcmp(argv[i], "--prompt")) prompt = next("--prompt");
      else if (!strcmp(argv[i], "--ids")) ids_arg = next("--ids");
      else if (!strcmp(argv[i], "--max-new")) max_new = (size_t)std::stoul(next("--max-new"));
      else if (!strcmp(argv[i], "--batch")) batch = (size_t)std::stoul(next("--batch"));
      else if (!strcmp(argv[i], "--device")) device = next("--device");
      else if (!strcmp(argv[i], "--precision")) precision = next("--precision");
      else if (!strcmp(argv[i], "--graph")) graph = true;
      else if (!strcmp(argv[i], "--bench")) {
        bench_b = (size_t)std::stoul(next("--bench"));
        bench_l = (size_t)std::stoul(next("--bench"));
        bench_steps = (size_t)std::stoul(next("--bench"));
      } else if (!strcmp(argv[i], "--help")) {
        std::cout << "usage: gpt2_inference [--model NAME] [--model-path FILE] [--vocab vocab.bin] [--prompt TEXT | --ids a,b,c] [--max-new N]\n"
                     "                      [--batch B] [--device CPU|GPU:i] [--precision bf16|fp8] [--graph] [--bench B L STEPS]\n";
        return 0;
      }
    }
    ExampleModels::register_defaults();
    auto& dm = DeviceManager::getInstance();
    const Device dev = device.empty() ? (dm.has_gpu() ? dm.getGPU(0) : dm.getHost()) : dm.getDevice(device);
    std::cout << "[gpt2_inference] model " << model << " on " << dev.name() << ", precision " << precision << std::endl;
    Graph g = load_or_create_model(model, model_path, dev, 1234);
    Layer* seq = g.get_layers().at(0);
    if (bench_b) {
      GPT2Engine eng(*seq, bench_b, 0, precision);
      const double pf = eng.time_prefill_ms(bench_b, bench_l, 5);
      const size_t ctx = std::min(bench_l, eng.max_len() - bench_steps - 2);
      const double dc = eng.time_decode_ms(bench_b, ctx, bench_steps, dev.is_gpu());
      std::cout << "{\"model\": \"" << model << "\", \"batch\": " << bench_b << ", \"seq_len\": " << bench_l << ", \"precision\": \"" << precision
                << "\", \"prefill_ms\": " << pf << ", \"prefill_tok_s\": " << (double)(bench_b * bench_l) / pf * 1e3 << ", \"decode_ms_per_step\": " << dc
                << ", \"decode_tok_s\": " << (double)bench_b / dc * 1e3 << ", \"decode_context\": " << ctx << ", \"host\": \"cpp\"}" << std::endl;
      const std::vector<double> prof = eng.persistent_decode_profile();
      if (!prof.empty()) {
        const char* names[5] = {"ln1+qkv", "attention", "out_proj", "ln2+fc1", "fc2"};
        double sum[5] = {0, 0, 0, 0, 0};
        const size_t nl = eng.num_layers();
        for (size_t l = 0; l < nl; ++l)
          for (int k = 0; k < 5; ++k) sum[k] += prof[1 + l * 5 + k];
        std::cout << "[persistent decode, CTA 0, last step] per layer (us):";
        for (int k = 0; k < 5; ++k) std::cout << ' ' << names[k] << ' ' << sum[k] / nl / 1e3;
        std::cout << "; embed " << prof[0] / 1e3 << " us, head " << prof[1 + nl * 5] / 1e3 << " us, argmax " << (prof.size() > nl * 5 + 2 ? prof[nl * 5 + 2] / 1e3 : 0.0) << " us" << std::endl;
      }
      return 0;
    }
    Tokenizer tok;
    const bool have_vocab = !vocab.empty();
    if (have_vocab) tok.load(vocab);
    std::vector<int32_t> ids;
    if (!ids_arg.empty()) {
      std::stringstream ss(ids_arg);
      std::string part;
      while (std::getline(ss, part, ',')) ids.push_back((int32_t)std::stol(part));
    } else if (have_vocab) {
      ids = tok.encode(prompt.empty() ? std::string("Hello, my name is") : prompt);
    } else {
      ids = {15496, 11, 616, 1438, 318};             // "Hello, my name is" in the GPT-2 vocabulary
    }
    if (ids.empty()) throw std::runtime_error("empty prompt");
    const size_t L = ids.size();
    std::vector<int32_t> batch_ids;
    for (size_t b = 0; b < batch; ++b) batch_ids.insert(batch_ids.end(), ids.begin(), ids.end());
    GPT2Engine eng(*seq, batch, 0, precision);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int32_t> out;
    size_t n = 0;
    if (graph) {
      std::vector<int32_t> first = eng.prefill(batch_ids, batch, L);
      std::vector<int32_t> rest = eng.decode_graph(first, max_new > 0 ? max_new - 1 : 0);
      n = 1 + (rest.size() / batch);
      out.resize(batch * n);
      for (size_t b = 0; b < batch; ++b) {
        out[b * n] = first[b];
        for (size_t s = 1; s < n; ++s) out[b * n + s] = rest[b * (n - 1) + s - 1];
      }
    } else {
      out = eng.generate(batch_ids, batch, L, max_new, 50256, &n);
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::vector<int32_t> row(out.begin(), out.begin() + (long)n);
    if (have_vocab) std::cout << (prompt.empty() ? tok.decode(ids) : prompt) << tok.decode(row) << std::endl;
    else {
      std::cout << "generated ids:";
      for (int32_t t : row) std::cout << ' ' << t;
      std::cout << std::endl;
    }
    std::cout << "[gpt2_inference] " << n << " tokens x " << batch << " sequence(s) in " << ms << " ms (" << (double)(n * batch) / ms * 1e3
              << " tok/s incl. prefill of " << L << ")" << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[gpt2_inference] error: " << e.what() << std::endl;
    return 1;
  }
}
