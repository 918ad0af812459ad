// TORCH_LIBRARY registration of the neighbour list over the C ABI, so that a TorchScript module can call it after
// torch.ops.load_library("libphysicsml_b200_torch.so") — the route the reference's OpenMM plugin needs: its graph module
// is scripted and run from OpenMM's C++ thread (plugins/openmm/physicsml_potential.py:130-140) and calls
// construct_edge_indices_and_attrs on every step (plugins/openmm/openmm_graph.py:148-156).
//
//   physicsml_b200_ts::radius_graph(Tensor positions, Tensor? cell, float cutoff, bool self_interaction)
//       -> (Tensor edge_index int64 [2,E], Tensor cell_shift_vector [E,3])
//
// Same semantics as physicsml_b200.neighbour_list.construct_edge_indices_and_attrs without bond features: edges sorted by
// (sender, receiver, shift); full PBC when a cell is given, open boundaries otherwise.  CUDA only, no CPU kernel is
// registered (a CPU tensor raises from the dispatcher).  Plain host code: all device work goes through the C ABI of
// libphysicsml_b200.so on the current stream.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>

#include "../../../include/physicsml_b200.h"

namespace {

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "physicsml_b200: ", what, " failed with code ", rc);
}

std::tuple<at::Tensor, at::Tensor> radius_graph(const at::Tensor& positions, const c10::optional<at::Tensor>& cell, double cutoff,
                                                bool self_interaction) {
  TORCH_CHECK(positions.is_cuda() && positions.dim() == 2 && positions.size(1) == 3, "positions must be a CUDA tensor [N, 3]");
  const c10::cuda::CUDAGuard guard(positions.device());
  auto stream = reinterpret_cast<pml_stream_t>(c10::cuda::getCurrentCUDAStream().stream());
  const at::Tensor pos = positions.detach().to(at::kFloat).contiguous();
  const int N = (int)pos.size(0);
  const auto i32 = pos.options().dtype(at::kInt);
  const float rc2 = (float)(cutoff * cutoff);
  at::Tensor deg = at::empty({N}, i32), rowptr = at::empty({N + 1}, i32);
  at::Tensor gptr = at::zeros({2}, i32);
  gptr.select(0, 1).fill_(N);
  at::Tensor err = at::zeros({2}, i32);
  const bool periodic = cell.has_value() && cell->defined();
  at::Tensor wpos, fl, bins, scratch, binptr, order, graphs;
  if (periodic) {
    const at::Tensor cells = cell->detach().to(pos.device(), at::kFloat).reshape({3, 3}).contiguous();
    graphs = at::empty({PML_NL_GRAPH_BYTES}, pos.options().dtype(at::kByte));
    at::Tensor nbins = at::zeros({1}, i32);
    check(pml_nl_setup(cells.data_ptr<float>(), 0, gptr.data_ptr<int>(), 1, (float)cutoff, graphs.data_ptr(), nbins.data_ptr<int>(),
                       err.data_ptr<int>() + 1, stream), "pml_nl_setup");
    const int max_bins = 64 + 4 * N;
    wpos = at::empty({N, 3}, pos.options());
    fl = at::empty({N, 3}, i32);
    bins = at::empty({N}, i32);
    scratch = at::zeros({2, max_bins}, i32);
    binptr = at::empty({max_bins + 1}, i32);
    order = at::empty({N}, i32);
    check(pml_nl_pbc_bin(pos.data_ptr<float>(), nullptr, graphs.data_ptr(), wpos.data_ptr<float>(), fl.data_ptr<int>(),
                         bins.data_ptr<int>(), scratch.data_ptr<int>(), N, stream), "pml_nl_pbc_bin");
    check(pml_scan_i32(scratch.data_ptr<int>(), binptr.data_ptr<int>(), max_bins, stream), "pml_scan_i32");
    check(pml_nl_pbc_fill_bins(bins.data_ptr<int>(), binptr.data_ptr<int>(), scratch.data_ptr<int>() + max_bins, order.data_ptr<int>(),
                               N, stream), "pml_nl_pbc_fill_bins");
    check(pml_nl_pbc_count(wpos.data_ptr<float>(), nullptr, bins.data_ptr<int>(), binptr.data_ptr<int>(), order.data_ptr<int>(),
                           graphs.data_ptr(), rc2, self_interaction ? 1 : 0, deg.data_ptr<int>(), N, stream), "pml_nl_pbc_count");
  } else {
    check(pml_nl_open_count(pos.data_ptr<float>(), nullptr, gptr.data_ptr<int>(), rc2, self_interaction ? 1 : 0, deg.data_ptr<int>(), N,
                            stream), "pml_nl_open_count");
  }
  check(pml_scan_i32(deg.data_ptr<int>(), rowptr.data_ptr<int>(), N, stream), "pml_scan_i32");
  const int E = N > 0 ? rowptr.select(0, N).item<int>() : 0;  // the one host sync: sizes the outputs
  at::Tensor keys = at::empty({E > 0 ? E : 1}, pos.options().dtype(at::kLong));
  auto* kp = reinterpret_cast<unsigned long long*>(keys.data_ptr<int64_t>());
  if (periodic) {
    check(pml_nl_pbc_fill(wpos.data_ptr<float>(), nullptr, bins.data_ptr<int>(), binptr.data_ptr<int>(), order.data_ptr<int>(),
                          graphs.data_ptr(), rc2, self_interaction ? 1 : 0, rowptr.data_ptr<int>(), kp, N, stream), "pml_nl_pbc_fill");
    check(pml_nl_sort_segments(kp, rowptr.data_ptr<int>(), N, err.data_ptr<int>(), stream), "pml_nl_sort_segments");
  } else {
    check(pml_nl_open_fill(pos.data_ptr<float>(), nullptr, gptr.data_ptr<int>(), rc2, self_interaction ? 1 : 0, rowptr.data_ptr<int>(), kp,
                           N, stream), "pml_nl_open_fill");
  }
  at::Tensor edge_index = at::empty({2, E}, pos.options().dtype(at::kLong));
  at::Tensor shift = at::zeros({E, 3}, pos.options());
  at::Tensor snd = at::empty({E}, i32), rcv = at::empty({E}, i32), perm = at::empty({E}, i32);
  check(pml_nl_emit(kp, rowptr.data_ptr<int>(), periodic ? fl.data_ptr<int>() : nullptr, reinterpret_cast<long long*>(edge_index.data_ptr<int64_t>()),
                    periodic ? shift.data_ptr<float>() : nullptr, snd.data_ptr<int>(), rcv.data_ptr<int>(), perm.data_ptr<int>(), N, E,
                    stream), "pml_nl_emit");
  if (periodic) {
    const at::Tensor flags = err.cpu();
    TORCH_CHECK(flags[0].item<int>() == 0, "an atom has more than 2048 neighbours");
    TORCH_CHECK((flags[1].item<int>() & 2) == 0, "cutoff spans more than 127 periodic images");
  }
  return std::make_tuple(edge_index, shift.to(positions.scalar_type()));
}

}  // namespace

TORCH_LIBRARY(physicsml_b200_ts, m) {
  m.def("radius_graph(Tensor positions, Tensor? cell, float cutoff, bool self_interaction) -> (Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(physicsml_b200_ts, CUDA, m) { m.impl("radius_graph", radius_graph); }
