// TEST INFRASTRUCTURE (oracle/_ref build only) — never part of the product path.
//
// Registers the reference's OWN custom all-reduce entry points (kernels/all_reduce/custom_all_reduce.cu, compiled where
// it lies by oracle/build_ref_cuda.py; prototypes kernels/ops.h:77-94) under torch.ops._ref_cuda_ar_C.*, for the N > 1
// legs of bench.py's `ref_cuda` arm. A separate library from _ref_cuda_C.so because this translation unit needs the
// driver API (cuPointerGetAttribute -> libcuda.so.1), which only exists on a GPU box.
#include <torch/library.h>
#include <torch/all.h>

#include "ops.h"   // /root/reference/kernels/ops.h

TORCH_LIBRARY(_ref_cuda_ar_C, m) {
  m.def("init_custom_ar", &init_custom_ar);
  m.def("all_reduce_reg", &all_reduce_reg);
  m.def("all_reduce_unreg", &all_reduce_unreg);
  m.def("dispose", &dispose);
  m.def("meta_size", &meta_size);
  m.def("register_buffer", &register_buffer);
  m.def("get_graph_buffer_ipc_meta", &get_graph_buffer_ipc_meta);
  m.def("register_graph_buffers", &register_graph_buffers);
}
