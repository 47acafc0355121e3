// dali_b200/host/pipeline.cc -- registries, OpSpec/OpSchema plumbing, TensorList storage and the per-sample batched
// executor (one stream per pipeline / GPU; operators run in insertion = topological order).
//
// Reference behaviour mirrored here:
//   InstantiateOperator: lookup by spec.SchemaName() in the registry selected by the "device" argument
//                        (dali/pipeline/operator/operator.cc:157-170)
//   OpTask::SetupOp / RunOp: fresh outputs, op->Setup(descs, ws); Resize when it returns true; op->Run(ws)
//                        (dali/pipeline/executor/executor2/exec_node_task.cc:252-346)
// The reference's prefetch queues, stream assignment and thread-pool scheduling (exec2) are out of scope
// (SURVEY.md 2.1 row 5): the hot-path ops are GPU-only and enqueue on one stream without host syncs.
#include "dali.h"
#include <nvtx3/nvToolsExt.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <mutex>

namespace dali {

struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ---------------------------------------------------------------------------------------------- TensorList storage
template <>
void TensorList<CPUBackend>::Free() {
  if (data_) cudaFreeHost(data_);
  data_ = nullptr; capacity_ = 0;
}
template <>
void TensorList<GPUBackend>::Free() {
  if (data_) cudaFree(data_);
  data_ = nullptr; capacity_ = 0;
}

template <typename Backend>
static void ResizeImpl(TensorListShape &shape_, DALIDataType &type_, std::vector<void *> &ptrs_, void *&data_, size_t &capacity_,
                       bool &owned_, const TensorListShape &shape, DALIDataType type, bool gpu) {
  shape_ = shape; type_ = type;
  const size_t esz = TypeSize(type);
  const int n = shape.num_samples();
  // every sample starts on a 256-byte boundary: kernels may use vector accesses relative to the sample base
  size_t total = 0;
  std::vector<size_t> offs(n);
  for (int i = 0; i < n; i++) { offs[i] = total; total += (static_cast<size_t>(shape.tensor_size(i)) * esz + 255) / 256 * 256; }
  if (total > capacity_) {
    if (data_) { if (gpu) cudaFree(data_); else cudaFreeHost(data_); }
    data_ = nullptr; capacity_ = 0;
    const size_t ncap = total + total / 8 + 256;
    if (gpu) CUDA_CALL(cudaMalloc(&data_, ncap)); else CUDA_CALL(cudaMallocHost(&data_, ncap));
    capacity_ = ncap;
  }
  ptrs_.resize(n);
  for (int i = 0; i < n; i++) ptrs_[i] = static_cast<uint8_t *>(data_) + offs[i];
  owned_ = true;
}
template <>
void TensorList<CPUBackend>::Resize(const TensorListShape &shape, DALIDataType type) {
  ResizeImpl<CPUBackend>(shape_, type_, ptrs_, data_, capacity_, owned_, shape, type, false);
}
template <>
void TensorList<GPUBackend>::Resize(const TensorListShape &shape, DALIDataType type) {
  ResizeImpl<GPUBackend>(shape_, type_, ptrs_, data_, capacity_, owned_, shape, type, true);
}

// ---------------------------------------------------------------------------------------------- registries
static std::map<std::string, std::unique_ptr<OpSchema>> &Schemas() {
  static std::map<std::string, std::unique_ptr<OpSchema>> s;
  return s;
}
OpSchema &SchemaRegistry::RegisterSchema(const std::string &name) {
  auto &m = Schemas();
  DALI_ENFORCE(!m.count(name), "OpSchema already registered for operator '", name, "'");
  m[name] = std::make_unique<OpSchema>(name);
  // arguments every operator has (op_schema.cc: "device", "num_threads", "max_batch_size", "seed", "bytes_per_sample_hint")
  m[name]->AddOptionalArg("device", "cpu | gpu | mixed", std::string("cpu"));
  m[name]->AddOptionalArg("num_threads", "", 1);
  m[name]->AddOptionalArg("max_batch_size", "", 1);
  m[name]->AddOptionalArg("seed", "", -1);
  m[name]->AddOptionalArg("bytes_per_sample_hint", "", std::vector<int64_t>{0});
  m[name]->AddOptionalArg("preserve", "", false);
  return *m[name];
}
const OpSchema *SchemaRegistry::TryGetSchema(const std::string &name) {
  auto &m = Schemas();
  auto it = m.find(name);
  return it == m.end() ? nullptr : it->second.get();
}
const OpSchema &SchemaRegistry::GetSchema(const std::string &name) {
  auto *s = TryGetSchema(name);
  DALI_ENFORCE(s, "Schema for operator '", name, "' not registered");
  return *s;
}
std::vector<std::string> SchemaRegistry::Names() {
  std::vector<std::string> r;
  for (auto &kv : Schemas()) r.push_back(kv.first);
  return r;
}

void OpSchema::CheckArgs(const OpSpec &spec) const {
  DALI_ENFORCE(spec.NumInput() >= min_in_ && spec.NumInput() <= max_in_, "Operator '", name_, "' takes ", min_in_, "..", max_in_,
               " inputs, got ", spec.NumInput());
  for (auto &r : required_)
    DALI_ENFORCE(spec.ArgumentDefined(r), "Argument '", r, "' is required by operator '", name_, "' but was not specified");
  for (auto &kv : spec.ArgumentInputs())
    DALI_ENFORCE(HasArgument(kv.first) && TensorArgAllowed(kv.first), "Argument '", kv.first, "' of operator '", name_,
                 "' does not accept a tensor (per-sample) input");
}

std::map<std::string, OperatorCreator> &OperatorRegistry::Registry(const std::string &backend) {
  static std::map<std::string, std::map<std::string, OperatorCreator>> r;
  return r[backend];
}
void OperatorRegistry::Register(const std::string &backend, const std::string &name, OperatorCreator c) {
  auto &r = Registry(backend);
  DALI_ENFORCE(!r.count(name), "Operator '", name, "' already registered for backend ", backend);   // operator_factory.h:55-58
  r[name] = std::move(c);
}
std::vector<std::string> OperatorRegistry::RegisteredNames(const std::string &backend) {
  std::vector<std::string> n;
  for (auto &kv : Registry(backend)) n.push_back(kv.first);
  return n;
}

std::unique_ptr<OperatorBase> InstantiateOperator(const OpSpec &spec) {
  const std::string device = spec.GetArgument<std::string>("device");
  auto &reg = OperatorRegistry::Registry(device);
  auto it = reg.find(spec.SchemaName());
  if (it == reg.end()) {
    DALI_FAIL(make_string("Operator '", spec.SchemaName(), "' is not registered for the '", device, "' backend. dali_b200 implements the "
                          "hot-path operators for the GPU / mixed backends only (there is no CPU fallback)."));
  }
  return it->second(spec);
}

// ---------------------------------------------------------------------------------------------- OpSpec
const ArgValue *OpSpec::FindArg(const std::string &n) const {
  auto it = args_.find(n);
  if (it != args_.end()) return &it->second;
  if (auto *s = SchemaRegistry::TryGetSchema(name_)) {
    auto d = s->Defaults().find(n);
    if (d != s->Defaults().end()) return &d->second;
  }
  return nullptr;
}

static double ArgAsDouble(const ArgValue &a, const std::string &n) {
  switch (a.kind) {
    case ArgValue::INT: case ArgValue::BOOL: return static_cast<double>(a.i);
    case ArgValue::FLOAT: return a.f;
    case ArgValue::FLOAT_VEC: if (a.fv.size() == 1) return a.fv[0]; break;
    case ArgValue::INT_VEC: if (a.iv.size() == 1) return static_cast<double>(a.iv[0]); break;
    default: break;
  }
  DALI_FAIL(make_string("Argument '", n, "' is not a scalar number"));
}

static double TensorArgScalar(const OpSpec &spec, const std::string &n, const Workspace *ws, int idx) {
  const auto &tl = ws->ArgumentInput(n);
  DALI_ENFORCE(idx < tl.num_samples(), "Argument input '", n, "' has ", tl.num_samples(), " samples, sample ", idx, " requested");
  DALI_ENFORCE(tl.shape().tensor_size(idx) == 1, "Argument input '", n, "' must hold one scalar per sample");
  const void *p = tl.raw_tensor(idx);
  switch (tl.type()) {
    case DALI_FLOAT: return *static_cast<const float *>(p);
    case DALI_FLOAT64: return *static_cast<const double *>(p);
    case DALI_INT32: return *static_cast<const int32_t *>(p);
    case DALI_INT64: return static_cast<double>(*static_cast<const int64_t *>(p));
    case DALI_UINT8: case DALI_BOOL: return *static_cast<const uint8_t *>(p);
    case DALI_INT16: return *static_cast<const int16_t *>(p);
    default: DALI_FAIL(make_string("Argument input '", n, "' has an unsupported type ", static_cast<int>(tl.type())));
  }
}

template <> double OpSpec::GetArgument<double>(const std::string &n, const Workspace *ws, int idx) const {
  if (ws && HasTensorArgument(n)) return TensorArgScalar(*this, n, ws, idx);
  auto *a = FindArg(n);
  DALI_ENFORCE(a, "Argument '", n, "' is not defined for operator '", name_, "'");
  return ArgAsDouble(*a, n);
}
template <> float OpSpec::GetArgument<float>(const std::string &n, const Workspace *ws, int idx) const {
  return static_cast<float>(GetArgument<double>(n, ws, idx));
}
template <> int OpSpec::GetArgument<int>(const std::string &n, const Workspace *ws, int idx) const {
  return static_cast<int>(GetArgument<double>(n, ws, idx));
}
template <> int64_t OpSpec::GetArgument<int64_t>(const std::string &n, const Workspace *ws, int idx) const {
  return static_cast<int64_t>(GetArgument<double>(n, ws, idx));
}
template <> bool OpSpec::GetArgument<bool>(const std::string &n, const Workspace *ws, int idx) const {
  return GetArgument<double>(n, ws, idx) != 0;
}
template <> DALIDataType OpSpec::GetArgument<DALIDataType>(const std::string &n, const Workspace *ws, int idx) const {
  return static_cast<DALIDataType>(GetArgument<int>(n, ws, idx));
}
template <> DALIInterpType OpSpec::GetArgument<DALIInterpType>(const std::string &n, const Workspace *ws, int idx) const {
  return static_cast<DALIInterpType>(GetArgument<int>(n, ws, idx));
}
template <> DALIImageType OpSpec::GetArgument<DALIImageType>(const std::string &n, const Workspace *ws, int idx) const {
  return static_cast<DALIImageType>(GetArgument<int>(n, ws, idx));
}
template <> std::string OpSpec::GetArgument<std::string>(const std::string &n, const Workspace *, int) const {
  auto *a = FindArg(n);
  DALI_ENFORCE(a, "Argument '", n, "' is not defined for operator '", name_, "'");
  DALI_ENFORCE(a->kind == ArgValue::STRING, "Argument '", n, "' is not a string");
  return a->s;
}
template <> TensorLayout OpSpec::GetArgument<TensorLayout>(const std::string &n, const Workspace *ws, int idx) const {
  return TensorLayout(GetArgument<std::string>(n, ws, idx));
}

#define DALI_TRY_GET(T)                                                                       \
  template <> bool OpSpec::TryGetArgument<T>(T & out, const std::string &n) const {           \
    if (!FindArg(n)) return false;                                                             \
    out = GetArgument<T>(n);                                                                   \
    return true;                                                                               \
  }
DALI_TRY_GET(float) DALI_TRY_GET(int) DALI_TRY_GET(bool) DALI_TRY_GET(std::string) DALI_TRY_GET(DALIDataType)
#undef DALI_TRY_GET
template <> bool OpSpec::TryGetArgument<std::vector<float>>(std::vector<float> &out, const std::string &n) const {
  if (!FindArg(n)) return false;
  out = GetRepeatedArgument<float>(n);
  return true;
}

template <> std::vector<float> OpSpec::GetRepeatedArgument<float>(const std::string &n) const {
  auto *a = FindArg(n);
  DALI_ENFORCE(a, "Argument '", n, "' is not defined for operator '", name_, "'");
  switch (a->kind) {
    case ArgValue::FLOAT_VEC: return a->fv;
    case ArgValue::INT_VEC: return std::vector<float>(a->iv.begin(), a->iv.end());
    case ArgValue::FLOAT: return { static_cast<float>(a->f) };
    case ArgValue::INT: case ArgValue::BOOL: return { static_cast<float>(a->i) };
    default: DALI_FAIL(make_string("Argument '", n, "' is not a list of numbers"));
  }
}
template <> std::vector<int> OpSpec::GetRepeatedArgument<int>(const std::string &n) const {
  auto f = GetRepeatedArgument<float>(n);
  return std::vector<int>(f.begin(), f.end());
}

std::vector<float> OpSpec::GetFloatVecArgument(const std::string &n, const Workspace *ws, int idx, int expected) const {
  std::vector<float> v;
  if (ws && HasTensorArgument(n)) {
    const auto &tl = ws->ArgumentInput(n);
    DALI_ENFORCE(idx < tl.num_samples(), "Argument input '", n, "': sample index out of range");
    const int64_t cnt = tl.shape().tensor_size(idx);
    v.resize(cnt);
    const void *p = tl.raw_tensor(idx);
    for (int64_t k = 0; k < cnt; k++) {
      switch (tl.type()) {
        case DALI_FLOAT: v[k] = static_cast<const float *>(p)[k]; break;
        case DALI_FLOAT64: v[k] = static_cast<float>(static_cast<const double *>(p)[k]); break;
        case DALI_INT32: v[k] = static_cast<float>(static_cast<const int32_t *>(p)[k]); break;
        case DALI_INT64: v[k] = static_cast<float>(static_cast<const int64_t *>(p)[k]); break;
        default: DALI_FAIL(make_string("Argument input '", n, "' has an unsupported type"));
      }
    }
  } else {
    v = GetRepeatedArgument<float>(n);
  }
  if (expected > 0) {
    if (static_cast<int>(v.size()) == 1 && expected > 1) v.assign(expected, v[0]);
    DALI_ENFORCE(static_cast<int>(v.size()) == expected, "Argument '", n, "' must have ", expected, " elements, got ", v.size());
  }
  return v;
}

// ---------------------------------------------------------------------------------------------- Pipeline
Pipeline::Pipeline(int max_batch_size, int num_threads, int device_id)
    : max_batch_size_(max_batch_size), num_threads_(num_threads), device_id_(device_id) {
  DALI_ENFORCE(max_batch_size > 0, "max_batch_size must be positive");
  if (device_id_ >= 0) {
    CUDA_CALL(cudaSetDevice(device_id_));
    CUDA_CALL(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  }
}

Pipeline::~Pipeline() {
  if (stream_) { cudaStreamSynchronize(stream_); }
  nodes_.clear();
  edges_.clear();
  if (stream_) cudaStreamDestroy(stream_);
}

Pipeline::Edge &Pipeline::GetEdge(const std::string &name, const std::string &device) {
  auto it = edges_.find(name + "/" + device);
  DALI_ENFORCE(it != edges_.end(), "Data node '", name, "' on device '", device, "' is not produced by any operator in the pipeline");
  return it->second;
}

void Pipeline::AddExternalInput(const std::string &name, const std::string &device, const std::string &layout) {
  DALI_ENFORCE(!built_, "Cannot add inputs after Build()");
  DALI_ENFORCE(device == "cpu" || device == "gpu", "external_source device must be 'cpu' or 'gpu'");
  Edge e;
  e.device = device; e.external = true; e.layout = layout;
  if (device == "cpu") e.cpu = std::make_unique<TensorList<CPUBackend>>(); else e.gpu = std::make_unique<TensorList<GPUBackend>>();
  DALI_ENFORCE(!edges_.count(name + "/" + device), "Duplicate data node name '", name, "'");
  edges_[name + "/" + device] = std::move(e);
}

void Pipeline::AddOperator(const OpSpec &spec_in, const std::string &inst_name) {
  DALI_ENFORCE(!built_, "Cannot add operators after Build()");
  OpSpec spec = spec_in;
  spec.AddArg("max_batch_size", MakeArg(max_batch_size_));
  spec.AddArg("num_threads", MakeArg(num_threads_));
  const auto &schema = SchemaRegistry::GetSchema(spec.SchemaName());
  schema.CheckArgs(spec);
  DALI_ENFORCE(spec.NumOutput() == schema.NumOutput(), "Operator '", spec.SchemaName(), "' produces ", schema.NumOutput(), " outputs");
  const std::string device = spec.GetArgument<std::string>("device");
  DALI_ENFORCE(device == "cpu" || device == "gpu" || device == "mixed", "Invalid device '", device, "'");
  DALI_ENFORCE(device == "cpu" || device_id_ >= 0, "Operator '", inst_name, "' needs a GPU but the pipeline was created with device_id=None");
  // inputs must exist already (operators are added in topological order)
  for (int i = 0; i < spec.NumInput(); i++) GetEdge(spec.Input(i).first, spec.Input(i).second);
  for (auto &kv : spec.ArgumentInputs()) GetEdge(kv.second, "cpu");
  for (int i = 0; i < spec.NumOutput(); i++) {
    Edge e;
    e.device = spec.Output(i).second;
    if (e.device == "cpu") e.cpu = std::make_unique<TensorList<CPUBackend>>(); else e.gpu = std::make_unique<TensorList<GPUBackend>>();
    const std::string key = spec.Output(i).first + "/" + e.device;
    DALI_ENFORCE(!edges_.count(key), "Duplicate data node name '", spec.Output(i).first, "'");
    edges_[key] = std::move(e);
  }
  Node n;
  n.spec = spec; n.name = inst_name;
  nodes_.push_back(std::move(n));
}

void Pipeline::SetOutputDescs(const std::vector<std::pair<std::string, std::string>> &outs) { output_names_ = outs; }

void Pipeline::Build() {
  DALI_ENFORCE(!built_, "Pipeline already built");
  for (auto &o : output_names_) GetEdge(o.first, o.second);
  if (device_id_ >= 0) CUDA_CALL(cudaSetDevice(device_id_));
  for (auto &n : nodes_) n.op = InstantiateOperator(n.spec);
  // decoder -> Resize fusion (resize straight from the decoder's Y / Cb / Cr planes, no RGB image in HBM): the decoded image must
  // have exactly one consumer and must not be a pipeline output.  Opt-in (DALIB200_FUSE_DECODE_RESIZE=1): it cuts the DRAM traffic
  // of the pair but measures ~4 % slower per batch than the two-kernel path on B200 (both are issue-bound).
  const char *fuse = getenv("DALIB200_FUSE_DECODE_RESIZE");
  if (fuse && fuse[0] == '1') {
    for (auto &c : nodes_) {
      auto *cons = dynamic_cast<PlanarConsumer *>(c.op.get());
      if (!cons || c.spec.NumInput() < 1) continue;
      const auto edge = c.spec.Input(0);
      int uses = 0;
      for (auto &o : nodes_) {
        for (int i = 0; i < o.spec.NumInput(); i++) uses += o.spec.Input(i) == edge;
      }
      for (auto &o : output_names_) uses += o == edge;
      if (uses != 1) continue;
      for (auto &pn : nodes_) {
        if (pn.spec.NumOutput() == 1 && pn.spec.Output(0) == edge) {
          if (auto *prod = dynamic_cast<PlanarProducer *>(pn.op.get())) { prod->EnableDeferredRun(); cons->AttachProducer(prod); }
        }
      }
    }
  }
  // Spectrogram -> MelFilterBank: one kernel when the spectrogram has no other consumer (the kernel pair is measurably slower and
  // writes + re-reads the spectrogram); DALIB200_NO_FUSION=1 keeps the operators separate
  if (!getenv("DALIB200_NO_FUSION")) {
    for (auto &c : nodes_) {
      auto *cons = dynamic_cast<SpectrumConsumer *>(c.op.get());
      if (!cons || c.spec.NumInput() < 1) continue;
      const auto edge = c.spec.Input(0);
      int uses = 0;
      for (auto &o : nodes_)
        for (int i = 0; i < o.spec.NumInput(); i++) uses += o.spec.Input(i) == edge;
      for (auto &o : output_names_) uses += o == edge;
      if (uses != 1) continue;
      for (auto &pn : nodes_) {
        if (pn.spec.NumOutput() == 1 && pn.spec.Output(0) == edge) {
          if (auto *prod = dynamic_cast<SpectrumProducer *>(pn.op.get())) { prod->EnableDeferredRun(); cons->AttachProducer(prod); }
        }
      }
    }
  }
  built_ = true;
}

void Pipeline::SetExternalInput(const std::string &name, const std::vector<const void *> &ptrs, const TensorListShape &shape,
                                DALIDataType type, const std::string &layout, bool no_copy) {
  DALI_ENFORCE(static_cast<int>(ptrs.size()) == shape.num_samples(), "SetExternalInput: pointer / shape count mismatch");
  DALI_ENFORCE(shape.num_samples() <= max_batch_size_, "External source batch (", shape.num_samples(), ") exceeds max_batch_size (",
               max_batch_size_, ")");
  Edge *e = nullptr;
  auto it = edges_.find(name + "/cpu");
  if (it == edges_.end()) it = edges_.find(name + "/gpu");
  DALI_ENFORCE(it != edges_.end() && it->second.external, "'", name, "' is not an external input of this pipeline");
  e = &it->second;
  const std::string &lay = layout.empty() ? e->layout : layout;
  if (e->device == "cpu") {
    // The data is borrowed for the duration of the iteration (the Python side keeps the buffers alive).
    std::vector<void *> p(ptrs.size());
    for (size_t i = 0; i < ptrs.size(); i++) p[i] = const_cast<void *>(ptrs[i]);
    e->cpu->ShareData(p, shape, type);
    e->cpu->set_stable(no_copy);
    e->cpu->SetLayout(lay);
  } else {
    e->gpu->Resize(shape, type);
    e->gpu->SetLayout(lay);
    const size_t esz = TypeSize(type);
    for (int i = 0; i < shape.num_samples(); i++) {
      const size_t bytes = static_cast<size_t>(shape.tensor_size(i)) * esz;
      if (bytes) CUDA_CALL(cudaMemcpyAsync(e->gpu->raw_mutable_tensor(i), ptrs[i], bytes, cudaMemcpyHostToDevice, stream_));
    }
  }
}

void Pipeline::Run() {
  DALI_ENFORCE(built_, "Pipeline must be built before Run()");
  if (device_id_ >= 0) CUDA_CALL(cudaSetDevice(device_id_));
  for (auto &n : nodes_) {
    Workspace ws;
    ws.set_stream(stream_);
    for (int i = 0; i < n.spec.NumInput(); i++) {
      Edge &e = GetEdge(n.spec.Input(i).first, n.spec.Input(i).second);
      if (e.cpu) ws.AddInput(e.cpu.get()); else ws.AddInput(e.gpu.get());
    }
    for (auto &kv : n.spec.ArgumentInputs()) ws.AddArgumentInput(kv.first, GetEdge(kv.second, "cpu").cpu.get());
    std::vector<Edge *> outs;
    for (int i = 0; i < n.spec.NumOutput(); i++) {
      Edge &e = GetEdge(n.spec.Output(i).first, n.spec.Output(i).second);
      outs.push_back(&e);
      if (e.cpu) ws.AddOutput(e.cpu.get()); else ws.AddOutput(e.gpu.get());
    }
    std::vector<OutputDesc> descs;
    static const bool timing = getenv("DALIB200_HOST_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    try {
      NvtxRange op_range(n.name.c_str());                  // like DomainTimeRange around Setup / Run (exec_node_task.cc:291,314)
      if (n.op->Setup(descs, ws)) {
        DALI_ENFORCE(descs.size() == outs.size(), "Operator returned ", descs.size(), " output descriptors for ", outs.size(), " outputs");
        for (size_t i = 0; i < outs.size(); i++) {
          if (outs[i]->cpu) outs[i]->cpu->Resize(descs[i].shape, descs[i].type); else outs[i]->gpu->Resize(descs[i].shape, descs[i].type);
        }
      }
      const auto t1 = std::chrono::steady_clock::now();
      n.op->Run(ws);
      if (timing) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[host timing] %-24s setup %.3f ms  run %.3f ms\n", n.spec.SchemaName().c_str(),
                std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
      }
    } catch (const std::exception &ex) {
      throw DALIException(make_string("Error in ", n.spec.GetArgument<std::string>("device"), " operator `", n.spec.SchemaName(),
                                      "` (", n.name, "): ", ex.what()));
    }
  }
}

bool Pipeline::OutputIsGPU(int i) const { return output_names_.at(i).second != "cpu"; }
const TensorList<CPUBackend> *Pipeline::OutputCPU(int i) const {
  auto it = edges_.find(output_names_.at(i).first + "/cpu");
  return it == edges_.end() ? nullptr : it->second.cpu.get();
}
const TensorList<GPUBackend> *Pipeline::OutputGPU(int i) const {
  auto it = edges_.find(output_names_.at(i).first + "/gpu");
  return it == edges_.end() ? nullptr : it->second.gpu.get();
}
void Pipeline::WaitOutputs() {
  if (stream_) CUDA_CALL(cudaStreamSynchronize(stream_));
  for (auto &n : nodes_) {
    try {
      n.op->CheckCompletion();
    } catch (const std::exception &ex) {
      throw DALIException(make_string("Error in ", n.spec.GetArgument<std::string>("device"), " operator `", n.spec.SchemaName(),
                                      "` (", n.name, "): ", ex.what()));
    }
  }
}

}  // namespace dali
