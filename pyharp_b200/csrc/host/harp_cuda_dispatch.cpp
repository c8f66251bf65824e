// harp_cuda_dispatch.cpp -- the definition pyharp declares and never provides:
//   template <int N> void harp::call_interpn_cuda(at::TensorIterator&, at::Tensor kdata, at::Tensor axis,
//                                                 at::Tensor dims, int nval)      (src/math/interpolation.hpp:12-14)
// The CUDA branches of the attenuators call it in comments only (src/opacity/rfm.cpp:160-161,190-191,
// s8_fuller.cpp:104-105, h2so4_simple.cpp:104-105); with this translation unit linked, un-commenting those four
// lines is all a maintainer needs for the reference's own RFM / S8Fuller / H2SO4Simple to run on the GPU.
// Same contract as call_interpn_cpu (src/math/math_dispatch.cpp:13-29): iter's output is (.., nval), its input the
// coordinate tensor (.., ndim) (possibly an expanded view); every point is interpolated by b200_interpn, which
// reproduces harp::interpn / harp::locate bit for bit.
#include <ATen/TensorIterator.h>
#include <torch/torch.h>

#include <b200opacity.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

namespace harp {

template <int N>
void call_interpn_cuda(at::TensorIterator& iter, at::Tensor kdata, at::Tensor axis, at::Tensor dims, int nval) {
  TORCH_CHECK(N >= nval, "N must be greater than or equal to nval");
  TORCH_CHECK(iter.dtype() == at::kDouble, "call_interpn_cuda: float64 only");
  auto out = iter.tensor(0);
  auto coord = iter.tensor(1);
  TORCH_CHECK(out.is_cuda() && kdata.is_cuda() && axis.is_cuda(), "call_interpn_cuda: CUDA tensors expected");
  const int ndim = (int)dims.size(0);
  auto hdims = dims.to(at::kCPU, at::kLong).contiguous();
  // points = all leading dimensions of the output; the iterator was declared with the last dimension squashed
  // (the aerosols pass a coordinate expanded over (ncol, nlyr, nprop): one coordinate per point, ndim = 1)
  const int64_t npoint = out.numel() / nval;
  at::Tensor c = coord.contiguous().to(out.device());
  if (c.numel() == out.numel() && ndim == 1 && nval > 1) c = c.view({npoint, nval}).select(1, 0).contiguous();
  TORCH_CHECK(c.numel() == npoint * ndim, "call_interpn_cuda: coordinate / output shape mismatch");
  at::Tensor o = out.is_contiguous() ? out : at::empty_like(out, at::MemoryFormat::Contiguous);
  // on the output's device and the caller's current stream of that device
  c10::cuda::CUDAGuard guard(out.device());
  int rc = b200_interpn(c.data_ptr<double>(), (long)npoint, ndim, kdata.contiguous().data_ptr<double>(),
                        axis.contiguous().data_ptr<double>(), hdims.data_ptr<int64_t>(), ndim, nval,
                        o.data_ptr<double>(), c10::cuda::getCurrentCUDAStream(out.device().index()).stream());
  TORCH_CHECK(rc == 0, "b200_interpn failed with code ", rc);
  if (!out.is_contiguous()) out.copy_(o);
}

template void call_interpn_cuda<1>(at::TensorIterator&, at::Tensor, at::Tensor, at::Tensor, int);
template void call_interpn_cuda<2>(at::TensorIterator&, at::Tensor, at::Tensor, at::Tensor, int);
template void call_interpn_cuda<3>(at::TensorIterator&, at::Tensor, at::Tensor, at::Tensor, int);

}  // namespace harp
