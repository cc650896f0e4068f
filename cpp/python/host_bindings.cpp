// tnn_b200._host: thin pybind11 layer over libtnn_b200 (the C++ framework).  NO torch here: arrays cross as numpy.
// Python tests drive the C++ layers / graphs / trainer through this module and compare them with torch oracles.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "kernels/simt_ops.cuh"
#include "tnn/bench.hpp"
#include "tnn/example_models.hpp"
#include "tnn/stepstate.hpp"
#include "tnn/inference.hpp"
#include "tnn/train.hpp"

namespace py = pybind11;
using namespace tnn;

static Device parse_device(const std::string& s) { return DeviceManager::getInstance().getDevice(s); }
// numpy (float32 / int32) -> host tensor
static Tensor from_numpy(const py::array& a) {
  py::array c = py::array::ensure(a, py::array::c_style | py::array::forcecast);
  Shape shape;
  for (py::ssize_t i = 0; i < c.ndim(); ++i) shape.push_back((size_t)c.shape(i));
  DType_t dt;
  if (c.dtype().is(py::dtype::of<float>())) dt = DType_t::FP32;
  else if (c.dtype().is(py::dtype::of<int32_t>())) dt = DType_t::INT32_T;
  else if (c.dtype().is(py::dtype::of<int64_t>())) dt = DType_t::INT64_T;
  else if (c.dtype().is(py::dtype::of<double>())) dt = DType_t::FP64;
  else throw std::invalid_argument("from_numpy: float32 / float64 / int32 / int64 arrays");
  Tensor t = make_tensor(dt, shape, getHost());
  std::memcpy(t->data(), c.data(), t->nbytes());
  if (dt == DType_t::FP64) return dtype_cast(*t, DType_t::FP32);
  if (dt == DType_t::INT64_T) return dtype_cast(*t, DType_t::INT32_T);
  return t;
}
// any tensor -> numpy float32 (or int32)
static py::array to_numpy(const Tensor& t) {
  if (!t) return py::array();
  std::vector<py::ssize_t> shape(t->shape().begin(), t->shape().end());
  if (t->data_type() == DType_t::INT32_T) {
    Tensor h = t->is_gpu() ? t->to_host() : t;
    py::array_t<int32_t> out(shape);
    std::memcpy(out.mutable_data(), h->data(), h->nbytes());
    return std::move(out);
  }
  const auto v = to_float_vector(*t);
  py::array_t<float> out(shape);
  std::memcpy(out.mutable_data(), v.data(), v.size() * 4);
  return std::move(out);
}

struct PyLayer {
  std::unique_ptr<Layer> layer;
  Device dev = getHost();
  void sync() const {
    if (dev.is_gpu()) {
      side_flow::join(dev);
      current_flow(dev).synchronize();
    }
  }
  Tensor put(const py::array& a) const {
    Tensor h = from_numpy(a);
    return dev.is_gpu() ? h->to_device(dev) : h;
  }
};

struct PyTrainer {
  std::unique_ptr<Graph> graph;
  std::unique_ptr<Loss> loss;
  std::unique_ptr<Optimizer> opt;
  std::unique_ptr<Scheduler> sched;
  std::unique_ptr<TrainStep> step;
};

struct PyGPT2 {
  std::unique_ptr<Layer> model;
  std::unique_ptr<GPT2Engine> engine;
};

PYBIND11_MODULE(_host, m) {
  m.doc() = "tnn_b200 C++ host (libtnn_b200) -- numpy in / numpy out";
  m.def("num_gpus", [] { return DeviceManager::getInstance().num_gpus(); });
  m.def("registered_layer_types", [] { return LayerFactory::registered_types(); });
  m.def("available_models", [] { return ExampleModels::available_models(); });
  m.def("model_config", [](const std::string& name) { return ExampleModels::create(name)->get_config().dump(-1); });
  m.def("launch_count", [] { return (unsigned long long)g_launch_count.load(); });
  // the flagship benchmark, in process (bench.py --host cpp); returns the JSON line as a string
  m.def("bench_train", [](const std::string& model, size_t batch, int steps, int warmup, int gpu, bool use_graph) {
    Json j;
    {
      py::gil_scoped_release rel;
      j = bench_train(model, batch, steps, warmup, gpu, use_graph);
    }
    return j.dump();
  }, py::arg("model") = "cifar100_wrn16_8", py::arg("batch") = 256, py::arg("steps") = 30, py::arg("warmup") = 5, py::arg("gpu") = 0,
     py::arg("use_graph") = true);

  py::class_<PyLayer>(m, "Layer")
      .def(py::init([](const std::string& config_json, const std::string& device, uint64_t seed) {
             auto p = std::make_unique<PyLayer>();
             p->layer = LayerFactory::create(TConfig::from_string(config_json));
             p->dev = parse_device(device);
             p->layer->init(p->dev, seed);
             return p;
           }), py::arg("config_json"), py::arg("device") = "CPU", py::arg("seed") = 1)
      .def("type", [](PyLayer& l) { return l.layer->type(); })
      .def("config", [](PyLayer& l) { return l.layer->get_config().dump(-1); })
      .def("set_training", [](PyLayer& l, bool f) { l.layer->set_training(f); })
      .def("output_shape", [](PyLayer& l, const Shape& s) { return l.layer->output_shape(s); })
      .def("forward", [](PyLayer& l, const py::array& x, size_t mb) {
        Tensor y = l.layer->forward(l.put(x), mb);
        l.sync();
        return to_numpy(y);
      }, py::arg("x"), py::arg("mb_id") = 0)
      .def("forward_multi", [](PyLayer& l, const std::vector<py::array>& xs, size_t mb) {
        Vec<Tensor> in;
        for (auto& x : xs) in.push_back(l.put(x));
        Vec<Tensor> out = l.layer->forward(in, mb);
        l.sync();
        std::vector<py::array> r;
        for (auto& t : out) r.push_back(to_numpy(t));
        return r;
      }, py::arg("xs"), py::arg("mb_id") = 0)
      .def("backward", [](PyLayer& l, const py::array& dy, size_t mb) {
        Tensor dx = l.layer->backward(l.put(dy), mb);
        l.sync();
        return to_numpy(dx);
      }, py::arg("dy"), py::arg("mb_id") = 0)
      // every parameter / gradient of the layer tree as {"layer_name.param": array}
      .def("params", [](PyLayer& l) {
        py::dict d;
        for (Layer* s : l.layer->all_layers())
          for (auto& p : s->params()) d[py::str(s->name() + "." + p.first)] = to_numpy(p.second);
        return d;
      })
      .def("grads", [](PyLayer& l) {
        l.sync();
        py::dict d;
        for (Layer* s : l.layer->all_layers())
          for (auto& p : s->grads()) d[py::str(s->name() + "." + p.first)] = to_numpy(p.second);
        return d;
      })
      .def("set_param", [](PyLayer& l, const std::string& qualified, const py::array& v) {
        for (Layer* s : l.layer->all_layers())
          for (auto& p : s->params())
            if (s->name() + "." + p.first == qualified) {
              Tensor h = from_numpy(v);
              if (h->size() != p.second->size()) throw std::invalid_argument("set_param: size mismatch for " + qualified);
              p.second->copy_from(*h->reshape(p.second->shape()));
              s->mark_params_dirty();
              return;
            }
        throw std::out_of_range("set_param: no parameter " + qualified);
      })
      .def("zero_grads", [](PyLayer& l) {
        for (Layer* s : l.layer->all_layers())
          for (auto& g : s->grads()) g.second->zero();
      });

  py::class_<PyTrainer>(m, "Trainer")
      .def(py::init([](const std::string& model, const std::string& device, uint64_t seed, const std::string& optimizer_json, const std::string& loss,
                       bool use_graph, const std::string& checkpoint) {
             auto t = std::make_unique<PyTrainer>();
             const Device dev = parse_device(device);
             t->opt = OptimizerFactory::create_from_config(TConfig::from_string(optimizer_json));
             if (!checkpoint.empty()) t->graph = std::make_unique<Graph>(load_checkpoint(checkpoint, dev, t->opt.get()));
             else t->graph = std::make_unique<Graph>(load_or_create_model(model, "", dev, seed));
             t->loss = LossFactory::create(loss);
             t->step = std::make_unique<TrainStep>(*t->graph, *t->loss, *t->opt, nullptr, 1, use_graph);
             return t;
           }), py::arg("model"), py::arg("device") = "CPU", py::arg("seed") = 1,
           py::arg("optimizer_json") = R"({"name":"adam","type":"adam","parameters":{"learning_rate":{"type":"double","value":0.001}}})",
           py::arg("loss") = "logsoftmax_crossentropy", py::arg("use_graph") = true, py::arg("checkpoint") = "")
      .def("step", [](PyTrainer& t, const py::array& x, const py::array& y) {
        Tensor hx = from_numpy(x), hy = from_numpy(y);
        Tensor st = (*t.step)(hx, hy);
        float h[2];
        st->copy_to_host(h, 8);
        return std::make_pair(h[0], h[1]);
      })
      .def("num_parameters", [](PyTrainer& t) { return t.graph->context().num_parameters(); })
      .def("captured", [](PyTrainer& t) { return t.step->captured(); })
      .def("param_slab", [](PyTrainer& t) { return to_numpy(t.graph->context().param_slab()); })
      .def("save", [](PyTrainer& t, const std::string& path) { save_checkpoint(*t.graph, path, t.opt.get(), nullptr, t.opt->step_count); })
      .def("graph_config", [](PyTrainer& t) { return t.graph->get_config().dump(-1); })
      .def("predict", [](PyTrainer& t, const py::array& x) {
        t.graph->set_training(false);
        GraphExecutor ex(*t.graph);
        Tensor h = from_numpy(x);
        Tensor out = ex.forward(t.graph->device().is_gpu() ? h->to_device(t.graph->device()) : h, 0);
        py::array r = to_numpy(out);
        t.graph->set_training(true);
        return r;
      });
  // GPT-2 inference engine of the C++ host (tnn/inference.hpp): int32 token arrays in, int32 / float32 arrays out
  py::class_<PyGPT2>(m, "GPT2Engine")
      .def(py::init([](const std::string& model_or_config, const std::string& device, uint64_t seed, size_t max_batch, size_t max_len,
                       const std::string& precision) {
             auto e = std::make_unique<PyGPT2>();
             const Device dev = parse_device(device);
             ExampleModels::register_defaults();
             if (!model_or_config.empty() && model_or_config[0] == '{') e->model = LayerFactory::create(TConfig::from_string(model_or_config));
             else e->model = ExampleModels::create(model_or_config);
             e->model->init(dev, seed);
             e->engine = std::make_unique<GPT2Engine>(*e->model, max_batch, max_len, precision);
             return e;
           }), py::arg("model") = "gpt2_small", py::arg("device") = "CPU", py::arg("seed") = 1, py::arg("max_batch") = 1, py::arg("max_len") = 0,
           py::arg("precision") = "bf16")
      .def("prefill", [](PyGPT2& e, const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& ids, bool want_logits) -> py::tuple {
        if (ids.ndim() != 2) throw std::invalid_argument("prefill: ids must be [B, L]");
        const size_t B = (size_t)ids.shape(0), L = (size_t)ids.shape(1);
        std::vector<int32_t> v(ids.data(), ids.data() + B * L);
        std::vector<float> lg;
        std::vector<int32_t> nx;
        {
          py::gil_scoped_release rel;
          nx = e.engine->prefill(v, B, L, want_logits ? &lg : nullptr);
        }
        py::array_t<int32_t> next((py::ssize_t)B);
        std::memcpy(next.mutable_data(), nx.data(), B * 4);
        if (!want_logits) return py::make_tuple(next, py::none());
        py::array_t<float> logits({(py::ssize_t)B, (py::ssize_t)e.engine->vocab()});
        std::memcpy(logits.mutable_data(), lg.data(), lg.size() * 4);
        return py::make_tuple(next, logits);
      }, py::arg("ids"), py::arg("return_logits") = false)
      .def("decode", [](PyGPT2& e, const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& tokens, bool want_logits) -> py::tuple {
        std::vector<int32_t> v(tokens.data(), tokens.data() + tokens.size());
        std::vector<float> lg;
        std::vector<int32_t> nx = e.engine->decode(v, want_logits ? &lg : nullptr);
        py::array_t<int32_t> next((py::ssize_t)nx.size());
        std::memcpy(next.mutable_data(), nx.data(), nx.size() * 4);
        if (!want_logits) return py::make_tuple(next, py::none());
        py::array_t<float> logits({(py::ssize_t)nx.size(), (py::ssize_t)e.engine->vocab()});
        std::memcpy(logits.mutable_data(), lg.data(), lg.size() * 4);
        return py::make_tuple(next, logits);
      }, py::arg("tokens"), py::arg("return_logits") = false)
      .def("decode_graph", [](PyGPT2& e, const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& tokens, size_t steps) {
        std::vector<int32_t> v(tokens.data(), tokens.data() + tokens.size());
        std::vector<int32_t> out;
        {
          py::gil_scoped_release rel;
          out = e.engine->decode_graph(v, steps);
        }
        const size_t B = v.size(), n = B ? out.size() / B : 0;
        py::array_t<int32_t> r({(py::ssize_t)B, (py::ssize_t)n});
        std::memcpy(r.mutable_data(), out.data(), out.size() * 4);
        return r;
      })
      .def("generate", [](PyGPT2& e, const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& ids, size_t max_new, long eos) {
        const size_t B = (size_t)ids.shape(0), L = (size_t)ids.shape(1);
        std::vector<int32_t> v(ids.data(), ids.data() + B * L);
        size_t n = 0;
        std::vector<int32_t> out = e.engine->generate(v, B, L, max_new, eos, &n);
        py::array_t<int32_t> r({(py::ssize_t)B, (py::ssize_t)n});
        std::memcpy(r.mutable_data(), out.data(), out.size() * 4);
        return r;
      }, py::arg("ids"), py::arg("max_new_tokens") = 50, py::arg("eos_token") = 50256)
      .def("full_forward_logits", [](PyGPT2& e, const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& ids) {
        const size_t B = (size_t)ids.shape(0), L = (size_t)ids.shape(1);
        std::vector<int32_t> v(ids.data(), ids.data() + B * L);
        std::vector<float> lg = e.engine->full_forward_logits(v, B, L);
        py::array_t<float> r({(py::ssize_t)B, (py::ssize_t)L, (py::ssize_t)(lg.size() / (B * L))});
        std::memcpy(r.mutable_data(), lg.data(), lg.size() * 4);
        return r;
      })
      .def("last_logits", [](PyGPT2& e) {
        std::vector<float> lg = e.engine->last_logits();
        const size_t V = e.engine->vocab();
        py::array_t<float> r({(py::ssize_t)(lg.size() / V), (py::ssize_t)V});
        std::memcpy(r.mutable_data(), lg.data(), lg.size() * 4);
        return r;
      })
      .def("uses_persistent_decode", [](PyGPT2& e) { return e.engine->uses_persistent_decode(); })
      .def("uses_fp8_kv_cache", [](PyGPT2& e) { return e.engine->uses_fp8_kv_cache(); })
      .def("enable_fp8", [](PyGPT2& e) { e.engine->enable_fp8(); })
      .def("precision", [](PyGPT2& e) { return e.engine->precision(); })
      .def("cur_len", [](PyGPT2& e) { return e.engine->cur_len(); })
      .def("max_len", [](PyGPT2& e) { return e.engine->max_len(); })
      .def("vocab", [](PyGPT2& e) { return e.engine->vocab(); })
      .def("num_layers", [](PyGPT2& e) { return e.engine->num_layers(); })
      .def("time_prefill_ms", [](PyGPT2& e, size_t B, size_t L, int iters, int warmup) {
        py::gil_scoped_release rel;
        return e.engine->time_prefill_ms(B, L, iters, warmup);
      }, py::arg("batch"), py::arg("seq_len"), py::arg("iters") = 5, py::arg("warmup") = 2)
      .def("time_decode_ms", [](PyGPT2& e, size_t B, size_t ctx, size_t steps, bool graph) {
        py::gil_scoped_release rel;
        return e.engine->time_decode_ms(B, ctx, steps, graph);
      }, py::arg("batch"), py::arg("context"), py::arg("steps") = 16, py::arg("graph") = true)
      .def("params", [](PyGPT2& e) {
        py::dict d;
        for (Layer* s : e.model->all_layers())
          for (auto& p : s->params()) d[py::str(s->name() + "." + p.first)] = to_numpy(p.second);
        return d;
      })
      .def("set_param", [](PyGPT2& e, const std::string& qualified, const py::array& v) {
        for (Layer* s : e.model->all_layers())
          for (auto& p : s->params())
            if (s->name() + "." + p.first == qualified) {
              Tensor h = from_numpy(v);
              if (h->size() != p.second->size()) throw std::invalid_argument("set_param: size mismatch for " + qualified);
              p.second->copy_from(*h->reshape(p.second->shape()));
              s->mark_params_dirty();
              return;
            }
        throw std::out_of_range("set_param: no parameter " + qualified);
      });
}
