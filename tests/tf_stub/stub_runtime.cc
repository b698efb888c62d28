// Driver side of the TensorFlow stub (tests only): owns the op / kernel registries the shim's static registrars fill
// and exposes a small C API so tests/test_tf_shim.py can list the registered signatures and run an OpKernel's
// Compute() on device pointers it supplies (inputs) and allocates on request (outputs / temps).
#include <cstring>

#include "tensorflow/core/framework/stub_core.h"

namespace tensorflow {
std::vector<StubOpDef> &StubOpRegistry() { static std::vector<StubOpDef> r; return r; }
std::vector<StubKernelEntry> &StubKernelRegistry() { static std::vector<StubKernelEntry> r; return r; }
}  // namespace tensorflow

using namespace tensorflow;

static std::string g_sig;

static const StubOpDef *find_op(const char *name) {
  for (auto &d : StubOpRegistry()) if (d.name == name) return &d;
  return nullptr;
}
static DataType type_of(const std::string &spec) { return spec.find("int32") != std::string::npos ? DT_INT32 : DT_FLOAT; }

extern "C" int tfstub_num_ops() { return (int)StubOpRegistry().size(); }
// "Name|attr;attr|input;input|output;output|shape_fn(0/1)"
extern "C" const char *tfstub_op_signature(int i) {
  const StubOpDef &d = StubOpRegistry().at(i);
  auto join = [](const std::vector<std::string> &v) { std::string s; for (size_t k = 0; k < v.size(); ++k) { if (k) s += ";"; s += v[k]; } return s; };
  g_sig = d.name + "|" + join(d.attrs) + "|" + join(d.inputs) + "|" + join(d.outputs) + "|" + (d.shape_fn ? "1" : "0");
  return g_sig.c_str();
}
extern "C" int tfstub_has_kernel(const char *op, const char *device) {
  int n = 0;
  for (auto &k : StubKernelRegistry()) if (k.def.op == op && k.def.device == device) ++n;
  return n;
}
// Runs op `name` on the GPU kernel the shim registered.  in_dims: concatenated dims of all inputs.
// Returns 0 = OK, 1 = the kernel (or its constructor) reported a TF error status (message in err), 2 = no such op /
// kernel, 3 = the shape function disagrees with the shapes the kernel allocated.
extern "C" int tfstub_run(const char *name, int n_attr, const char **attr_names, const int *attr_vals, int n_in,
                          void **in_ptrs, const int *in_ranks, const long long *in_dims, StubAlloc alloc, char *err,
                          int errlen) {
  auto fail = [&](int rc, const std::string &m) { if (err && errlen > 0) { std::strncpy(err, m.c_str(), errlen - 1); err[errlen - 1] = 0; } return rc; };
  const StubOpDef *def = find_op(name);
  if (!def) return fail(2, "op not registered");
  const StubKernelEntry *ke = nullptr;
  for (auto &k : StubKernelRegistry()) if (k.def.op == name && k.def.device == DEVICE_GPU) ke = &k;
  if (!ke) return fail(2, "no GPU kernel registered");
  if ((int)def->inputs.size() != n_in) return fail(2, "wrong number of inputs");
  OpKernelConstruction cons;
  for (int i = 0; i < n_attr; ++i) cons.attrs[attr_names[i]] = attr_vals[i];
  std::unique_ptr<OpKernel> k(ke->make(&cons));
  if (!cons.status.ok()) return fail(1, cons.status.error_message());
  OpKernelContext ctx;
  ctx.alloc = alloc;
  for (auto &o : def->outputs) ctx.out_types.push_back(type_of(o));
  shape_inference::InferenceContext ic;
  ic.attrs = cons.attrs;
  const long long *d = in_dims;
  for (int i = 0; i < n_in; ++i) {
    std::vector<int64_t> dims(d, d + in_ranks[i]);
    d += in_ranks[i];
    ctx.inputs.emplace_back(type_of(def->inputs[i]), TensorShape(dims), in_ptrs[i]);
    shape_inference::ShapeHandle sh; sh.dims = dims; sh.known = true;
    ic.inputs.push_back(sh);
  }
  k->Compute(&ctx);
  if (!ctx.status.ok()) return fail(1, ctx.status.error_message());
  if (ctx.outputs.size() != def->outputs.size()) return fail(3, "kernel did not allocate every declared output");
  if (def->shape_fn) {
    Status s = def->shape_fn(&ic);
    if (!s.ok()) return fail(3, "shape fn failed: " + s.error_message());
    for (size_t i = 0; i < def->outputs.size(); ++i) {
      if (i >= ic.outputs.size() || !ic.outputs[i].known) return fail(3, "shape fn left an output unset");
      if (ic.outputs[i].dims != ctx.outputs[(int)i].shape().vec()) return fail(3, "shape fn and kernel disagree on output " + std::to_string(i));
    }
  }
  return 0;
}
