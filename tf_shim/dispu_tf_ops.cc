// TensorFlow registration shim for libdispu_b200.so — SOURCE ONLY in this repository: it needs the
// TensorFlow headers (tf.sysconfig.get_compile_flags()), which are not installable in the build
// container.  Build it on a box that has the reference's TensorFlow with tf_shim/build_tf_ops.sh.
//
// It registers the SAME op names, inputs, outputs and attrs as the reference wrappers
// (tf_ops/sampling/tf_sampling.cpp:14-63, tf_ops/grouping/tf_grouping.cpp:14-64,
// tf_ops/interpolation/tf_interpolate.cpp:12-46, tf_ops/nn_distance/tf_nndistance.cpp:3-18), so
// tf_sampling.py / tf_grouping.py / tf_interpolate.py / tf_nndistance.py load it unchanged through
// tf.load_op_library, and gives every op a DEVICE_GPU kernel that forwards to the C ABI
// (include/dispu_b200.h).  New relative to the reference: ThreeNN / ThreeInterpolate(+Grad) run on
// the GPU (the reference registers them for DEVICE_CPU only, tf_interpolate.cpp:187,222,262).
// No DEVICE_CPU kernel is registered for any op: there is no CPU fallback.
//
// One .so per reference directory is produced from this single file by defining exactly one of
// DISPU_TF_SAMPLING / DISPU_TF_GROUPING / DISPU_TF_INTERPOLATE / DISPU_TF_NNDISTANCE (so each
// tf_*_so.so registers only its own ops, like the reference's).
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "../include/dispu_b200.h"

namespace {
using namespace tensorflow;
using shape_inference::InferenceContext;
using shape_inference::ShapeHandle;

// The reference launches on the legacy default stream (no stream argument anywhere, e.g.
// tf_sampling_g.cu:203-205); passing nullptr keeps that behaviour bit for bit.
inline dispu_stream_t Stream(OpKernelContext *) { return nullptr; }

inline Status Check(const char *op, int rc) {
  if (rc == DISPU_OK) return Status::OK();
  if (rc < 0) return errors::InvalidArgument(op, ": ", dispu_error_string(rc));
  return errors::Internal(op, ": CUDA error ", rc, " (", dispu_error_string(rc), ")");
}

#define DISPU_OUT(ctx, i, shape, var) \
  Tensor *var = nullptr;              \
  OP_REQUIRES_OK(ctx, ctx->allocate_output(i, TensorShape shape, &var))
#define F(t) ((t).flat<float>().data())
#define I(t) ((t).flat<int>().data())
}  // namespace

// ================================================================================== sampling
#ifdef DISPU_TF_SAMPLING
REGISTER_OP("ProbSample").Input("inp: float32").Input("inpr: float32").Output("out: int32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle r; TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 2, &r));
      c->set_output(0, r); return Status::OK(); });
REGISTER_OP("FarthestPointSample").Attr("npoint: int").Input("inp: float32").Output("out: int32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle s; TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &s));
      int npoint = 0; TF_RETURN_IF_ERROR(c->GetAttr("npoint", &npoint));
      c->set_output(0, c->MakeShape({c->Dim(s, 0), npoint})); return Status::OK(); });
REGISTER_OP("GatherPoint").Input("inp: float32").Input("idx: int32").Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle a, b; TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &a));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 2, &b));
      c->set_output(0, c->MakeShape({c->Dim(a, 0), c->Dim(b, 1), c->Dim(a, 2)})); return Status::OK(); });
REGISTER_OP("GatherPointGrad").Input("inp: float32").Input("idx: int32").Input("out_g: float32")
    .Output("inp_g: float32")
    .SetShapeFn([](InferenceContext *c) { c->set_output(0, c->input(0)); return Status::OK(); });

class FarthestPointSampleOp : public OpKernel {
 public:
  explicit FarthestPointSampleOp(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("npoint", &npoint_));
    OP_REQUIRES(c, npoint_ > 0, errors::InvalidArgument("FarthestPointSample expects positive npoint"));
  }
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0);
    OP_REQUIRES(ctx, inp.dims() == 3 && inp.dim_size(2) == 3,
                errors::InvalidArgument("FarthestPointSample expects (batch_size,num_points,3) inp shape"));
    const int b = inp.dim_size(0), n = inp.dim_size(1);
    DISPU_OUT(ctx, 0, ({b, npoint_}), out);
    Tensor temp;  // only touched beyond the on-chip capacity; same (32,n) scratch as the reference
    float *temp_ptr = nullptr;
    if (n > dispu_fps_onchip_capacity()) {
      OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_FLOAT, TensorShape{32, n}, &temp));
      temp_ptr = F(temp);
    }
    OP_REQUIRES_OK(ctx, Check("FarthestPointSample", dispu_fps(b, n, npoint_, F(inp), temp_ptr, I(*out), Stream(ctx))));
  }
 private:
  int npoint_;
};
REGISTER_KERNEL_BUILDER(Name("FarthestPointSample").Device(DEVICE_GPU), FarthestPointSampleOp);

class GatherPointOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0), &idx = ctx->input(1);
    OP_REQUIRES(ctx, inp.dims() == 3 && inp.dim_size(2) == 3,
                errors::InvalidArgument("GatherPoint expects (batch_size,num_points,3) inp shape"));
    const int b = inp.dim_size(0), n = inp.dim_size(1);
    OP_REQUIRES(ctx, idx.dims() == 2 && idx.dim_size(0) == b,
                errors::InvalidArgument("GatherPoint expects (batch_size,num_result) idx shape"));
    const int m = idx.dim_size(1);
    DISPU_OUT(ctx, 0, ({b, m, 3}), out);
    OP_REQUIRES_OK(ctx, Check("GatherPoint", dispu_gather_point(b, n, m, F(inp), I(idx), F(*out), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("GatherPoint").Device(DEVICE_GPU), GatherPointOp);

class GatherPointGradOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0), &idx = ctx->input(1), &out_g = ctx->input(2);
    OP_REQUIRES(ctx, inp.dims() == 3 && inp.dim_size(2) == 3,
                errors::InvalidArgument("GatherPointGradGpuOp expects (batch_size,num_points,3) inp"));
    const int b = inp.dim_size(0), n = inp.dim_size(1);
    OP_REQUIRES(ctx, idx.dims() == 2 && idx.dim_size(0) == b,
                errors::InvalidArgument("GatherPointGradGpuOp expects (batch_size,num_result) idx shape"));
    const int m = idx.dim_size(1);
    OP_REQUIRES(ctx, out_g.dims() == 3 && out_g.dim_size(0) == b && out_g.dim_size(1) == m && out_g.dim_size(2) == 3,
                errors::InvalidArgument("GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape"));
    DISPU_OUT(ctx, 0, ({b, n, 3}), inp_g);
    OP_REQUIRES_OK(ctx, Check("GatherPointGrad",
                              dispu_gather_point_grad(b, n, m, F(out_g), I(idx), F(*inp_g), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("GatherPointGrad").Device(DEVICE_GPU), GatherPointGradOp);

class ProbSampleOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0), &inpr = ctx->input(1);
    OP_REQUIRES(ctx, inp.dims() == 2, errors::InvalidArgument("ProbSample expects (batch_size,num_choices) inp shape"));
    const int b = inp.dim_size(0), n = inp.dim_size(1);
    OP_REQUIRES(ctx, inpr.dims() == 2 && inpr.dim_size(0) == b,
                errors::InvalidArgument("ProbSample expects (batch_size,num_points) inpr shape"));
    const int m = inpr.dim_size(1);
    DISPU_OUT(ctx, 0, ({b, m}), out);
    Tensor temp;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_FLOAT, TensorShape{b, n}, &temp));
    OP_REQUIRES_OK(ctx, Check("ProbSample", dispu_prob_sample(b, n, m, F(inp), F(inpr), F(temp), I(*out), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("ProbSample").Device(DEVICE_GPU), ProbSampleOp);
#endif  // DISPU_TF_SAMPLING

// ================================================================================== grouping
#ifdef DISPU_TF_GROUPING
REGISTER_OP("QueryBallPoint").Attr("nsample: int").Input("xyz1: float32").Input("xyz2: float32")
    .Input("radius: float32").Output("idx: int32").Output("pts_cnt: int32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle q; TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &q));
      int nsample = 0; TF_RETURN_IF_ERROR(c->GetAttr("nsample", &nsample));
      c->set_output(0, c->MakeShape({c->Dim(q, 0), c->Dim(q, 1), nsample}));
      c->set_output(1, c->MakeShape({c->Dim(q, 0), c->Dim(q, 1)})); return Status::OK(); });
REGISTER_OP("SelectionSort").Attr("k: int").Input("dist: float32").Output("outi: int32").Output("out: float32")
    .SetShapeFn([](InferenceContext *c) { c->set_output(0, c->input(0)); c->set_output(1, c->input(0)); return Status::OK(); });
REGISTER_OP("GroupPoint").Input("points: float32").Input("idx: int32").Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle p, i; TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &p));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &i));
      c->set_output(0, c->MakeShape({c->Dim(i, 0), c->Dim(i, 1), c->Dim(i, 2), c->Dim(p, 2)})); return Status::OK(); });
REGISTER_OP("GroupPointGrad").Input("points: float32").Input("idx: int32").Input("grad_out: float32")
    .Output("grad_points: float32")
    .SetShapeFn([](InferenceContext *c) { c->set_output(0, c->input(0)); return Status::OK(); });

class QueryBallPointOp : public OpKernel {
 public:
  explicit QueryBallPointOp(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("nsample", &nsample_));
    OP_REQUIRES(c, nsample_ > 0, errors::InvalidArgument("QueryBallPoint expects positive nsample"));
  }
  void Compute(OpKernelContext *ctx) override {
    const Tensor &xyz1 = ctx->input(0), &xyz2 = ctx->input(1), &radius = ctx->input(2);
    OP_REQUIRES(ctx, xyz1.dims() == 3 && xyz1.dim_size(2) == 3,
                errors::InvalidArgument("QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape."));
    const int b = xyz1.dim_size(0), n = xyz1.dim_size(1);
    OP_REQUIRES(ctx, xyz2.dims() == 3 && xyz2.dim_size(2) == 3 && xyz2.dim_size(0) == b,
                errors::InvalidArgument("QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape."));
    const int m = xyz2.dim_size(1);
    DISPU_OUT(ctx, 0, ({b, m, nsample_}), idx);
    DISPU_OUT(ctx, 1, ({b, m}), cnt);
    OP_REQUIRES_OK(ctx, Check("QueryBallPoint", dispu_query_ball_point(b, n, m, F(radius), nsample_, F(xyz1), F(xyz2),
                                                                       I(*idx), I(*cnt), Stream(ctx))));
  }
 private:
  int nsample_;
};
REGISTER_KERNEL_BUILDER(Name("QueryBallPoint").Device(DEVICE_GPU), QueryBallPointOp);

class SelectionSortOp : public OpKernel {
 public:
  explicit SelectionSortOp(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("k", &k_));
    OP_REQUIRES(c, k_ > 0, errors::InvalidArgument("SelectionSort expects positive k"));
  }
  void Compute(OpKernelContext *ctx) override {
    const Tensor &dist = ctx->input(0);
    OP_REQUIRES(ctx, dist.dims() == 3, errors::InvalidArgument("SelectionSort expects (b,m,n) dist shape."));
    const int b = dist.dim_size(0), m = dist.dim_size(1), n = dist.dim_size(2);
    DISPU_OUT(ctx, 0, ({b, m, n}), outi);
    DISPU_OUT(ctx, 1, ({b, m, n}), out);
    OP_REQUIRES_OK(ctx, Check("SelectionSort", dispu_selection_sort(b, n, m, k_, F(dist), I(*outi), F(*out), Stream(ctx))));
  }
 private:
  int k_;
};
REGISTER_KERNEL_BUILDER(Name("SelectionSort").Device(DEVICE_GPU), SelectionSortOp);

class GroupPointOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1);
    OP_REQUIRES(ctx, points.dims() == 3, errors::InvalidArgument("GroupPoint expects (batch_size, num_points, channel) points shape"));
    const int b = points.dim_size(0), n = points.dim_size(1), c = points.dim_size(2);
    OP_REQUIRES(ctx, idx.dims() == 3 && idx.dim_size(0) == b,
                errors::InvalidArgument("GroupPoint expects (batch_size, npoints, nsample) idx shape"));
    const int m = idx.dim_size(1), ns = idx.dim_size(2);
    DISPU_OUT(ctx, 0, ({b, m, ns, c}), out);
    OP_REQUIRES_OK(ctx, Check("GroupPoint", dispu_group_point(b, n, c, m, ns, F(points), I(idx), F(*out), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("GroupPoint").Device(DEVICE_GPU), GroupPointOp);

class GroupPointGradOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1), &grad_out = ctx->input(2);
    OP_REQUIRES(ctx, points.dims() == 3, errors::InvalidArgument("GroupPointGrad expects (batch_size, num_points, channel) points shape"));
    const int b = points.dim_size(0), n = points.dim_size(1), c = points.dim_size(2);
    OP_REQUIRES(ctx, idx.dims() == 3 && idx.dim_size(0) == b,
                errors::InvalidArgument("GroupPointGrad expects (batch_size, npoints, nsample) idx shape"));
    const int m = idx.dim_size(1), ns = idx.dim_size(2);
    OP_REQUIRES(ctx, grad_out.dims() == 4 && grad_out.dim_size(0) == b && grad_out.dim_size(1) == m &&
                         grad_out.dim_size(2) == ns && grad_out.dim_size(3) == c,
                errors::InvalidArgument("GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape"));
    DISPU_OUT(ctx, 0, ({b, n, c}), grad_points);
    OP_REQUIRES_OK(ctx, Check("GroupPointGrad", dispu_group_point_grad(b, n, c, m, ns, F(grad_out), I(idx),
                                                                       F(*grad_points), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("GroupPointGrad").Device(DEVICE_GPU), GroupPointGradOp);
#endif  // DISPU_TF_GROUPING

// ================================================================================== interpolation
#ifdef DISPU_TF_INTERPOLATE
REGISTER_OP("ThreeNN").Input("xyz1: float32").Input("xyz2: float32").Output("dist: float32").Output("idx: int32")
    .SetShapeFn([](InferenceContext *c) { c->set_output(0, c->input(0)); c->set_output(1, c->input(0)); return Status::OK(); });
REGISTER_OP("ThreeInterpolate").Input("points: float32").Input("idx: int32").Input("weight: float32")
    .Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle p, i; TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &p));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &i));
      c->set_output(0, c->MakeShape({c->Dim(p, 0), c->Dim(i, 1), c->Dim(p, 2)})); return Status::OK(); });
REGISTER_OP("ThreeInterpolateGrad").Input("points: float32").Input("idx: int32").Input("weight: float32")
    .Input("grad_out: float32").Output("grad_points: float32")
    .SetShapeFn([](InferenceContext *c) { c->set_output(0, c->input(0)); return Status::OK(); });

class ThreeNNOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &xyz1 = ctx->input(0), &xyz2 = ctx->input(1);
    OP_REQUIRES(ctx, xyz1.dims() == 3 && xyz1.dim_size(2) == 3, errors::InvalidArgument("ThreeNN expects (b,n,3) xyz1 shape"));
    const int b = xyz1.dim_size(0), n = xyz1.dim_size(1);
    OP_REQUIRES(ctx, xyz2.dims() == 3 && xyz2.dim_size(2) == 3 && xyz2.dim_size(0) == b,
                errors::InvalidArgument("ThreeNN expects (b,m,3) xyz2 shape"));
    const int m = xyz2.dim_size(1);
    DISPU_OUT(ctx, 0, ({b, n, 3}), dist);
    DISPU_OUT(ctx, 1, ({b, n, 3}), idx);
    OP_REQUIRES_OK(ctx, Check("ThreeNN", dispu_three_nn(b, n, m, F(xyz1), F(xyz2), F(*dist), I(*idx), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("ThreeNN").Device(DEVICE_GPU), ThreeNNOp);

class ThreeInterpolateOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1), &weight = ctx->input(2);
    OP_REQUIRES(ctx, points.dims() == 3, errors::InvalidArgument("ThreeInterpolate expects (b,m,c) points shape"));
    const int b = points.dim_size(0), m = points.dim_size(1), c = points.dim_size(2);
    OP_REQUIRES(ctx, idx.dims() == 3 && idx.dim_size(0) == b && idx.dim_size(2) == 3,
                errors::InvalidArgument("ThreeInterpolate expects (b,n,3) idx shape"));
    const int n = idx.dim_size(1);
    OP_REQUIRES(ctx, weight.dims() == 3 && weight.dim_size(0) == b && weight.dim_size(1) == n && weight.dim_size(2) == 3,
                errors::InvalidArgument("ThreeInterpolate expects (b,n,3) weight shape"));
    DISPU_OUT(ctx, 0, ({b, n, c}), out);
    OP_REQUIRES_OK(ctx, Check("ThreeInterpolate", dispu_three_interpolate(b, m, c, n, F(points), I(idx), F(weight),
                                                                          F(*out), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("ThreeInterpolate").Device(DEVICE_GPU), ThreeInterpolateOp);

class ThreeInterpolateGradOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1), &weight = ctx->input(2), &grad_out = ctx->input(3);
    OP_REQUIRES(ctx, points.dims() == 3, errors::InvalidArgument("ThreeInterpolateGrad expects (b,m,c) points shape"));
    const int b = points.dim_size(0), m = points.dim_size(1), c = points.dim_size(2);
    OP_REQUIRES(ctx, idx.dims() == 3 && idx.dim_size(0) == b, errors::InvalidArgument("ThreeInterpolateGrad expects (b,n,3) idx shape"));
    const int n = idx.dim_size(1);
    OP_REQUIRES(ctx, grad_out.dims() == 3 && grad_out.dim_size(0) == b && grad_out.dim_size(1) == n && grad_out.dim_size(2) == c,
                errors::InvalidArgument("ThreeInterpolateGrad expects (b,n,c) grad_out shape"));
    DISPU_OUT(ctx, 0, ({b, m, c}), grad_points);
    OP_REQUIRES_OK(ctx, Check("ThreeInterpolateGrad", dispu_three_interpolate_grad(b, n, c, m, F(grad_out), I(idx), F(weight),
                                                                                  F(*grad_points), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("ThreeInterpolateGrad").Device(DEVICE_GPU), ThreeInterpolateGradOp);
#endif  // DISPU_TF_INTERPOLATE

// ================================================================================== nn_distance
#ifdef DISPU_TF_NNDISTANCE
REGISTER_OP("NnDistance").Input("xyz1: float32").Input("xyz2: float32").Output("dist1: float32")
    .Output("idx1: int32").Output("dist2: float32").Output("idx2: int32");  // no shape fn, as in the reference
REGISTER_OP("NnDistanceGrad").Input("xyz1: float32").Input("xyz2: float32").Input("grad_dist1: float32")
    .Input("idx1: int32").Input("grad_dist2: float32").Input("idx2: int32").Output("grad_xyz1: float32")
    .Output("grad_xyz2: float32");

class NnDistanceOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &xyz1 = ctx->input(0), &xyz2 = ctx->input(1);
    OP_REQUIRES(ctx, xyz1.dims() == 3, errors::InvalidArgument("NnDistance requires xyz1 be of shape (batch,#points,3)"));
    OP_REQUIRES(ctx, xyz1.dim_size(2) == 3, errors::InvalidArgument("NnDistance only accepts 3d point set xyz1"));
    const int b = xyz1.dim_size(0), n = xyz1.dim_size(1);
    OP_REQUIRES(ctx, xyz2.dims() == 3, errors::InvalidArgument("NnDistance requires xyz2 be of shape (batch,#points,3)"));
    OP_REQUIRES(ctx, xyz2.dim_size(2) == 3, errors::InvalidArgument("NnDistance only accepts 3d point set xyz2"));
    OP_REQUIRES(ctx, xyz2.dim_size(0) == b, errors::InvalidArgument("NnDistance expects xyz1 and xyz2 have same batch size"));
    const int m = xyz2.dim_size(1);
    DISPU_OUT(ctx, 0, ({b, n}), dist1);
    DISPU_OUT(ctx, 1, ({b, n}), idx1);
    DISPU_OUT(ctx, 2, ({b, m}), dist2);
    DISPU_OUT(ctx, 3, ({b, m}), idx2);
    OP_REQUIRES_OK(ctx, Check("NnDistance", dispu_nn_distance(b, n, F(xyz1), m, F(xyz2), F(*dist1), I(*idx1), F(*dist2),
                                                              I(*idx2), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("NnDistance").Device(DEVICE_GPU), NnDistanceOp);

class NnDistanceGradOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext *ctx) override {
    const Tensor &xyz1 = ctx->input(0), &xyz2 = ctx->input(1), &g1 = ctx->input(2), &i1 = ctx->input(3),
                 &g2 = ctx->input(4), &i2 = ctx->input(5);
    OP_REQUIRES(ctx, xyz1.dims() == 3 && xyz1.dim_size(2) == 3, errors::InvalidArgument("NnDistanceGrad requires xyz1 be of shape (batch,#points,3)"));
    const int b = xyz1.dim_size(0), n = xyz1.dim_size(1);
    OP_REQUIRES(ctx, xyz2.dims() == 3 && xyz2.dim_size(2) == 3 && xyz2.dim_size(0) == b,
                errors::InvalidArgument("NnDistanceGrad requires xyz2 be of shape (batch,#points,3)"));
    const int m = xyz2.dim_size(1);
    OP_REQUIRES(ctx, g1.shape() == (TensorShape{b, n}) && i1.shape() == (TensorShape{b, n}),
                errors::InvalidArgument("NnDistanceGrad requires grad_dist1/idx1 be of shape(batch,#points)"));
    OP_REQUIRES(ctx, g2.shape() == (TensorShape{b, m}) && i2.shape() == (TensorShape{b, m}),
                errors::InvalidArgument("NnDistanceGrad requires grad_dist2/idx2 be of shape(batch,#points)"));
    DISPU_OUT(ctx, 0, ({b, n, 3}), o1);
    DISPU_OUT(ctx, 1, ({b, m, 3}), o2);
    OP_REQUIRES_OK(ctx, Check("NnDistanceGrad", dispu_nn_distance_grad(b, n, F(xyz1), m, F(xyz2), F(g1), I(i1), F(g2), I(i2),
                                                                       F(*o1), F(*o2), Stream(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("NnDistanceGrad").Device(DEVICE_GPU), NnDistanceGradOp);
#endif  // DISPU_TF_NNDISTANCE
