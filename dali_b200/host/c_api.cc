// dali_b200/host/c_api.cc -- C interface of the host layer (pipeline construction / run / outputs), bound from Python
// with ctypes (dali_b200/backend.py).  Style follows the reference's C API (include/dali/dali.h:43-164,428-1250:
// opaque handles, integer status, thread-local last error); only what the Python layer needs is exposed.
#include <algorithm>
#include <cstring>
#include <string>
#include "dali.h"

using namespace dali;  // NOLINT

static thread_local std::string g_err;

#define API_BEGIN try {
#define API_END                                       \
  }                                                   \
  catch (const std::exception &e) { g_err = e.what(); return 1; } \
  catch (...) { g_err = "unknown error"; return 1; }  \
  return 0;

template <typename TL>
static void FillInfo(const TL *tl, int *n, int *ndim, int *dtype, char *layout, int *contiguous) {
  DALI_ENFORCE(tl, "output not available");
  *n = tl->num_samples(); *ndim = tl->sample_dim(); *dtype = tl->type();
  snprintf(layout, 16, "%s", tl->GetLayout().str().c_str());
  *contiguous = tl->IsContiguous();
}

extern "C" {

const char *dalihLastError() { return g_err.c_str(); }

// ---- OpSpec
int dalihOpSpecCreate(void **spec, const char *schema) { API_BEGIN *spec = new OpSpec(schema); API_END }
int dalihOpSpecDestroy(void *spec) { API_BEGIN delete static_cast<OpSpec *>(spec); API_END }
int dalihOpSpecAddArgInt(void *s, const char *n, int64_t v) { API_BEGIN static_cast<OpSpec *>(s)->AddArg(n, MakeArg(v)); API_END }
int dalihOpSpecAddArgFloat(void *s, const char *n, double v) { API_BEGIN static_cast<OpSpec *>(s)->AddArg(n, MakeArg(v)); API_END }
int dalihOpSpecAddArgBool(void *s, const char *n, int v) { API_BEGIN static_cast<OpSpec *>(s)->AddArg(n, MakeArg(v != 0)); API_END }
int dalihOpSpecAddArgString(void *s, const char *n, const char *v) { API_BEGIN static_cast<OpSpec *>(s)->AddArg(n, MakeArg(std::string(v))); API_END }
int dalihOpSpecAddArgFloatVec(void *s, const char *n, const float *v, int cnt) {
  API_BEGIN static_cast<OpSpec *>(s)->AddArg(n, MakeArg(std::vector<float>(v, v + cnt))); API_END
}
int dalihOpSpecAddArgIntVec(void *s, const char *n, const int64_t *v, int cnt) {
  API_BEGIN static_cast<OpSpec *>(s)->AddArg(n, MakeArg(std::vector<int64_t>(v, v + cnt))); API_END
}
int dalihOpSpecAddInput(void *s, const char *name, const char *device) { API_BEGIN static_cast<OpSpec *>(s)->AddInput(name, device); API_END }
int dalihOpSpecAddOutput(void *s, const char *name, const char *device) { API_BEGIN static_cast<OpSpec *>(s)->AddOutput(name, device); API_END }
int dalihOpSpecAddArgumentInput(void *s, const char *arg, const char *input) { API_BEGIN static_cast<OpSpec *>(s)->AddArgumentInput(arg, input); API_END }

// ---- schema introspection (drives the generation of fn.* wrappers, like ops/__init__.py:553-715 in the reference)
int dalihNumSchemas() { return static_cast<int>(SchemaRegistry::Names().size()); }
int dalihSchemaName(int i, char *out, int cap) {
  API_BEGIN
  auto names = SchemaRegistry::Names();
  DALI_ENFORCE(i >= 0 && i < static_cast<int>(names.size()), "schema index out of range");
  snprintf(out, cap, "%s", names[i].c_str());
  API_END
}
// writes a '\n'-separated list: "name|tensor_ok(0/1)|required(0/1)"
int dalihSchemaArgs(const char *schema, char *out, int cap) {
  API_BEGIN
  const auto &s = SchemaRegistry::GetSchema(schema);
  std::string r;
  for (auto &n : s.ArgNames()) {
    const bool req = std::find(s.Required().begin(), s.Required().end(), n) != s.Required().end();
    r += n + "|" + (s.TensorArgAllowed(n) ? "1" : "0") + "|" + (req ? "1" : "0") + "\n";
  }
  DALI_ENFORCE(static_cast<int>(r.size()) < cap, "buffer too small");
  snprintf(out, cap, "%s", r.c_str());
  API_END
}
int dalihSchemaInfo(const char *schema, int *min_in, int *max_in, int *num_out, char *doc, int cap) {
  API_BEGIN
  const auto &s = SchemaRegistry::GetSchema(schema);
  *min_in = s.MinNumInput(); *max_in = s.MaxNumInput(); *num_out = s.NumOutput();
  if (doc) snprintf(doc, cap, "%s", s.doc().c_str());
  API_END
}
int dalihOperatorRegistered(const char *schema, const char *backend) {
  return OperatorRegistry::Registry(backend).count(schema) ? 1 : 0;
}

// ---- Pipeline
int dalihPipelineCreate(void **pipe, int max_batch, int num_threads, int device_id) {
  API_BEGIN *pipe = new Pipeline(max_batch, num_threads, device_id); API_END
}
int dalihPipelineDestroy(void *p) { API_BEGIN delete static_cast<Pipeline *>(p); API_END }
int dalihPipelineAddExternalInput(void *p, const char *name, const char *device, const char *layout) {
  API_BEGIN static_cast<Pipeline *>(p)->AddExternalInput(name, device, layout ? layout : ""); API_END
}
int dalihPipelineAddOperator(void *p, void *spec, const char *inst_name) {
  API_BEGIN static_cast<Pipeline *>(p)->AddOperator(*static_cast<OpSpec *>(spec), inst_name); API_END
}
int dalihPipelineSetOutputs(void *p, int n, const char *const *names, const char *const *devices) {
  API_BEGIN
  std::vector<std::pair<std::string, std::string>> o;
  for (int i = 0; i < n; i++) o.emplace_back(names[i], devices[i]);
  static_cast<Pipeline *>(p)->SetOutputDescs(o);
  API_END
}
int dalihPipelineBuild(void *p) { API_BEGIN static_cast<Pipeline *>(p)->Build(); API_END }
int dalihPipelineFeedInput(void *p, const char *name, int n, const void *const *ptrs, const int64_t *shapes, int ndim, int dtype,
                           const char *layout) {
  API_BEGIN
  TensorListShape sh(n, ndim);
  for (int i = 0; i < n; i++) sh.set_tensor_shape(i, TensorShape(shapes + static_cast<size_t>(i) * ndim, shapes + static_cast<size_t>(i + 1) * ndim));
  static_cast<Pipeline *>(p)->SetExternalInput(name, std::vector<const void *>(ptrs, ptrs + n), sh, static_cast<DALIDataType>(dtype),
                                               layout ? layout : "");
  API_END
}
// no_copy != 0: the caller keeps the buffers valid AND unmodified until the iteration has completed (external_source(no_copy=True))
int dalihPipelineFeedInputEx(void *p, const char *name, int n, const void *const *ptrs, const int64_t *shapes, int ndim, int dtype,
                             const char *layout, int no_copy) {
  API_BEGIN
  TensorListShape sh(n, ndim);
  for (int i = 0; i < n; i++) sh.set_tensor_shape(i, TensorShape(shapes + static_cast<size_t>(i) * ndim, shapes + static_cast<size_t>(i + 1) * ndim));
  static_cast<Pipeline *>(p)->SetExternalInput(name, std::vector<const void *>(ptrs, ptrs + n), sh, static_cast<DALIDataType>(dtype),
                                               layout ? layout : "", no_copy != 0);
  API_END
}
int dalihPipelineRun(void *p) { API_BEGIN static_cast<Pipeline *>(p)->Run(); API_END }
int dalihPipelineWait(void *p) { API_BEGIN static_cast<Pipeline *>(p)->WaitOutputs(); API_END }
int dalihPipelineNumOutputs(void *p) { return static_cast<Pipeline *>(p)->NumOutputs(); }
int dalihPipelineStream(void *p, void **stream) { API_BEGIN *stream = static_cast<Pipeline *>(p)->stream(); API_END }

int dalihPipelineOutputInfo(void *p, int i, int *is_gpu, int *n, int *ndim, int *dtype, char *layout, int *contiguous) {
  API_BEGIN
  auto *pp = static_cast<Pipeline *>(p);
  *is_gpu = pp->OutputIsGPU(i);
  if (*is_gpu) FillInfo(pp->OutputGPU(i), n, ndim, dtype, layout, contiguous); else FillInfo(pp->OutputCPU(i), n, ndim, dtype, layout, contiguous);
  API_END
}
int dalihPipelineOutputData(void *p, int i, int64_t *shapes, void **ptrs) {
  API_BEGIN
  auto *pp = static_cast<Pipeline *>(p);
  auto fill = [&](const auto *tl) {
    DALI_ENFORCE(tl, "output not available");
    const int nd = tl->sample_dim();
    for (int s = 0; s < tl->num_samples(); s++) {
      for (int d = 0; d < nd; d++) shapes[static_cast<size_t>(s) * nd + d] = tl->tensor_shape_span(s)[d];
      ptrs[s] = const_cast<void *>(tl->raw_tensor(s));
    }
  };
  if (pp->OutputIsGPU(i)) fill(pp->OutputGPU(i)); else fill(pp->OutputCPU(i));
  API_END
}

}  // extern "C"
