// kaldi-decoder_b200/csrc/python/module.cc
//
// pybind11 module `_kaldi_decoder`: the Python surface of the reference's
// hot-path classes (kaldi-decoder/python/csrc/{faster-decoder,decodable-ctc,
// decodable-itf,kaldi-decoder}.cc) with the same names, argument names and
// defaults, plus the FST value types the reference gets from kaldifst (absent
// here) and the additive batched decoder.

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "kaldi-decoder_b200/csrc/decodable-ctc.h"
#include "kaldi-decoder_b200/csrc/decodable-itf.h"
#include "kaldi-decoder_b200/csrc/faster-decoder.h"
#include "kaldi-decoder_b200/csrc/fst-io.h"
#include "kaldi-decoder_b200/csrc/lattice-faster-decoder.h"
#include "kaldi-decoder_b200/csrc/simple-decoder.h"
#include "kaldifst/csrc/remove-eps-local.h"
#include "kd_capi.h"

namespace py = pybind11;
using namespace kaldi_decoder;  // NOLINT

namespace {

using FloatArray = py::array_t<float, py::array::c_style | py::array::forcecast>;

// python/csrc/decodable-itf.cc:16-39 of the reference
class PyDecodableInterface : public DecodableInterface {
 public:
  using DecodableInterface::DecodableInterface;
  float LogLikelihood(int32_t frame, int32_t index) override {
    PYBIND11_OVERRIDE_PURE_NAME(float, DecodableInterface, "log_likelihood", LogLikelihood, frame,
                                index);
  }
  bool IsLastFrame(int32_t frame) const override {
    PYBIND11_OVERRIDE_PURE_NAME(bool, DecodableInterface, "is_last_frame", IsLastFrame, frame);
  }
  int32_t NumFramesReady() const override {
    PYBIND11_OVERRIDE_NAME(int32_t, DecodableInterface, "num_frames_ready", NumFramesReady);
  }
  int32_t NumIndices() const override {
    PYBIND11_OVERRIDE_PURE_NAME(int32_t, DecodableInterface, "num_indices", NumIndices);
  }
};

std::unique_ptr<DecodableCtc> MakeDecodableCtc(const FloatArray &feats, int32_t offset) {
  if (feats.ndim() != 2) throw std::runtime_error("DecodableCtc: feats must be 2-D");
  FloatMatrix m(feats.data(), static_cast<int32_t>(feats.shape(0)),
                static_cast<int32_t>(feats.shape(1)));
  return std::make_unique<DecodableCtc>(m, offset);
}

fst::StdVectorFst FstFromArrays(int32_t num_states, int32_t start,
                                const py::array_t<int64_t, py::array::c_style | py::array::forcecast> &row_off,
                                const py::array_t<int32_t, py::array::c_style | py::array::forcecast> &il,
                                const py::array_t<int32_t, py::array::c_style | py::array::forcecast> &ol,
                                const FloatArray &w,
                                const py::array_t<int32_t, py::array::c_style | py::array::forcecast> &ns,
                                const FloatArray &fin) {
  if (row_off.size() != num_states + 1 || fin.size() != num_states)
    throw std::runtime_error("from_arrays: row_offsets/final have the wrong length");
  fst::StdVectorFst f;
  f.ReserveStates(num_states);
  for (int32_t s = 0; s < num_states; ++s) f.AddState();
  const int64_t *off = row_off.data();
  for (int32_t s = 0; s < num_states; ++s) {
    f.SetFinal(s, fst::TropicalWeight(fin.data()[s]));
    f.ReserveArcs(s, static_cast<size_t>(off[s + 1] - off[s]));
    for (int64_t a = off[s]; a < off[s + 1]; ++a)
      f.AddArc(s, fst::StdArc(il.data()[a], ol.data()[a], fst::TropicalWeight(w.data()[a]),
                              ns.data()[a]));
  }
  if (start >= 0) f.SetStart(start);
  return f;
}

py::tuple FstToArrays(const fst::Fst<fst::StdArc> &f) {
  const int32_t n = fst::CountStates(f);
  int64_t e = 0;
  for (int32_t s = 0; s < n; ++s) e += static_cast<int64_t>(f.NumArcs(s));
  py::array_t<int64_t> off(n + 1);
  py::array_t<int32_t> il(e), ol(e), ns(e);
  py::array_t<float> w(e), fin(n);
  int64_t k = 0;
  off.mutable_data()[0] = 0;
  for (int32_t s = 0; s < n; ++s) {
    for (fst::ArcIterator<fst::Fst<fst::StdArc>> aiter(f, s); !aiter.Done(); aiter.Next(), ++k) {
      const fst::StdArc &arc = aiter.Value();
      il.mutable_data()[k] = arc.ilabel;
      ol.mutable_data()[k] = arc.olabel;
      w.mutable_data()[k] = arc.weight.Value();
      ns.mutable_data()[k] = arc.nextstate;
    }
    off.mutable_data()[s + 1] = k;
    fin.mutable_data()[s] = f.Final(s).Value();
  }
  return py::make_tuple(n, f.Start(), off, il, ol, w, ns, fin);
}

std::vector<std::tuple<int, int, float, int>> StdArcsOf(const fst::Fst<fst::StdArc> &f, int s) {
  std::vector<std::tuple<int, int, float, int>> out;
  for (fst::ArcIterator<fst::Fst<fst::StdArc>> aiter(f, s); !aiter.Done(); aiter.Next()) {
    const fst::StdArc &arc = aiter.Value();
    out.emplace_back(arc.ilabel, arc.olabel, arc.weight.Value(), arc.nextstate);
  }
  return out;
}

}  // namespace

PYBIND11_MODULE(_kaldi_decoder, m) {
  m.doc() = "B200-native kaldi-decoder: pybind11 binding";

  // ---- FST value types (stand-ins for kaldifst.StdVectorFst / kaldifst.Lattice)
  // The reference's constructors take kaldifst's Fst<StdArc> / VectorFst / ConstFst
  // (python/csrc/faster-decoder.cc:34-42): the same three types here.
  py::class_<fst::StdFst>(m, "StdFst")
      .def_property_readonly("start", &fst::StdFst::Start)
      .def_property_readonly("num_states",
                             [](const fst::StdFst &f) { return fst::CountStates(f); })
      .def("num_arcs", &fst::StdFst::NumArcs, py::arg("state"))
      .def("final", [](const fst::StdFst &f, int s) { return f.Final(s).Value(); },
           py::arg("state"))
      .def("arcs", &StdArcsOf, py::arg("state"),
           "(ilabel, olabel, weight, nextstate) of every arc leaving `state`")
      .def("to_arrays", &FstToArrays)
      .def("to_str", [](const fst::StdFst &f) { return WriteFstText(f); })
      .def("write", [](const fst::StdFst &f, const std::string &p) { WriteFst(f, p); },
           py::arg("filename"), "OpenFst binary, fst type \"vector\"")
      .def_property_readonly("fst_type", [](const fst::StdFst &f) { return f.Type(); })
      .def_property_readonly("content_id", &fst::StdFst::ContentId,
                             "process-unique number of the states and arcs this object holds "
                             "(copies share it, a modification replaces it); decoders built from "
                             "FSTs with the same content_id share one device graph");

  py::class_<fst::StdVectorFst, fst::StdFst>(m, "StdVectorFst")
      .def(py::init<>())
      .def(py::init([](const fst::StdFst &f) { return fst::StdVectorFst(f); }), py::arg("fst"))
      .def_static("from_arrays", &FstFromArrays, py::arg("num_states"), py::arg("start"),
                  py::arg("row_offsets"), py::arg("ilabel"), py::arg("olabel"), py::arg("weight"),
                  py::arg("nextstate"), py::arg("final"))
      .def_static("read", &ReadFst, py::arg("filename"))
      .def_static("from_str", &ReadFstText, py::arg("s"), py::arg("acceptor") = false)
      .def("add_state", &fst::StdVectorFst::AddState)
      .def("set_start", &fst::StdVectorFst::SetStart, py::arg("state"))
      .def("set_final",
           [](fst::StdVectorFst &f, int s, float w) { f.SetFinal(s, fst::StdArc::Weight(w)); },
           py::arg("state"), py::arg("weight") = 0.0f)
      .def("add_arc",
           [](fst::StdVectorFst &f, int s, int il, int ol, float w, int ns) {
             f.AddArc(s, fst::StdArc(il, ol, fst::StdArc::Weight(w), ns));
           },
           py::arg("state"), py::arg("ilabel"), py::arg("olabel"), py::arg("weight"),
           py::arg("nextstate"));

  // Immutable CSR form: what an OpenFst "const" file holds and what the device graph is
  // built from without an intermediate copy per state.
  py::class_<fst::StdConstFst, fst::StdFst>(m, "StdConstFst")
      .def(py::init<>())
      .def(py::init([](const fst::StdFst &f) { return fst::StdConstFst(f); }), py::arg("fst"))
      .def_static("read", [](const std::string &p) { return fst::StdConstFst(ReadFst(p)); },
                  py::arg("filename"))
      .def_static(
          "from_arrays",
          [](int32_t num_states, int32_t start,
             const py::array_t<int64_t, py::array::c_style | py::array::forcecast> &row_off,
             const py::array_t<int32_t, py::array::c_style | py::array::forcecast> &il,
             const py::array_t<int32_t, py::array::c_style | py::array::forcecast> &ol,
             const FloatArray &w,
             const py::array_t<int32_t, py::array::c_style | py::array::forcecast> &ns,
             const FloatArray &fin) {
            return fst::StdConstFst(FstFromArrays(num_states, start, row_off, il, ol, w, ns, fin));
          },
          py::arg("num_states"), py::arg("start"), py::arg("row_offsets"), py::arg("ilabel"),
          py::arg("olabel"), py::arg("weight"), py::arg("nextstate"), py::arg("final"));

  py::class_<fst::Lattice>(m, "Lattice")
      .def(py::init<>())
      .def_property_readonly("start", &fst::Lattice::Start)
      .def_property_readonly("num_states", &fst::Lattice::NumStates)
      .def("num_arcs", &fst::Lattice::NumArcs, py::arg("state"))
      .def("final",
           [](const fst::Lattice &f, int s) {
             return std::make_pair(f.Final(s).Value1(), f.Final(s).Value2());
           },
           py::arg("state"), "(graph, acoustic) final weight")
      .def("arcs",
           [](const fst::Lattice &f, int s) {
             std::vector<std::tuple<int, int, float, float, int>> out;
             for (fst::ArcIterator<fst::Lattice> aiter(f, s); !aiter.Done(); aiter.Next()) {
               const fst::LatticeArc &arc = aiter.Value();
               out.emplace_back(arc.ilabel, arc.olabel, arc.weight.Value1(), arc.weight.Value2(),
                                arc.nextstate);
             }
             return out;
           },
           py::arg("state"), "(ilabel, olabel, graph, acoustic, nextstate) of every arc");

  m.def(
      "get_linear_symbol_sequence",
      [](const fst::Lattice &lat) {
        std::vector<int32_t> isyms, osyms;
        fst::LatticeWeight tot;
        bool ok = GetLinearSymbolSequence(lat, &isyms, &osyms, &tot);
        return py::make_tuple(ok, isyms, osyms, std::make_pair(tot.Value1(), tot.Value2()));
      },
      py::arg("fst"),
      "(ok, isymbols_out, osymbols_out, (graph, acoustic) total weight) of a linear lattice");

  // ---- DecodableInterface / DecodableCtc
  py::class_<DecodableInterface, PyDecodableInterface>(m, "DecodableInterface")
      .def(py::init<>())
      .def("log_likelihood", &DecodableInterface::LogLikelihood, py::arg("frame"), py::arg("index"))
      .def("is_last_frame", &DecodableInterface::IsLastFrame, py::arg("frame"))
      .def("num_frames_ready", &DecodableInterface::NumFramesReady)
      .def("num_indices", &DecodableInterface::NumIndices);

  py::class_<DecodableCtc, DecodableInterface>(m, "DecodableCtc")
      .def(py::init(&MakeDecodableCtc), py::arg("feats"), py::arg("offset") = 0);

  // ---- FasterDecoderOptions / FasterDecoder
  py::class_<FasterDecoderOptions>(m, "FasterDecoderOptions")
      .def(py::init<float, int32_t, int32_t, float, float>(), py::arg("beam") = 16.0,
           py::arg("max_active") = std::numeric_limits<int32_t>::max(), py::arg("min_active") = 20,
           py::arg("beam_delta") = 0.5, py::arg("hash_ratio") = 2.0)
      .def_readwrite("beam", &FasterDecoderOptions::beam)
      .def_readwrite("max_active", &FasterDecoderOptions::max_active)
      .def_readwrite("min_active", &FasterDecoderOptions::min_active)
      .def_readwrite("beam_delta", &FasterDecoderOptions::beam_delta)
      .def_readwrite("hash_ratio", &FasterDecoderOptions::hash_ratio)
      .def("__str__", &FasterDecoderOptions::ToString);

  // lattice-faster-decoder.h:23-134 of the reference (the struct only; see that header)
  py::class_<LatticeFasterDecoderConfig>(m, "LatticeFasterDecoderConfig")
      .def(py::init<float, int32_t, int32_t, float, int32_t, bool, float, float, float, int32_t,
                    int32_t>(),
           py::arg("beam") = 16.0, py::arg("max_active") = std::numeric_limits<int32_t>::max(),
           py::arg("min_active") = 200, py::arg("lattice_beam") = 10.0,
           py::arg("prune_interval") = 25, py::arg("determinize_lattice") = true,
           py::arg("beam_delta") = 0.5, py::arg("hash_ratio") = 2.0, py::arg("prune_scale") = 0.1,
           py::arg("memory_pool_tokens_block_size") = 1 << 8,
           py::arg("memory_pool_links_block_size") = 1 << 8)
      .def_readwrite("beam", &LatticeFasterDecoderConfig::beam)
      .def_readwrite("max_active", &LatticeFasterDecoderConfig::max_active)
      .def_readwrite("min_active", &LatticeFasterDecoderConfig::min_active)
      .def_readwrite("lattice_beam", &LatticeFasterDecoderConfig::lattice_beam)
      .def_readwrite("prune_interval", &LatticeFasterDecoderConfig::prune_interval)
      .def_readwrite("determinize_lattice", &LatticeFasterDecoderConfig::determinize_lattice)
      .def_readwrite("beam_delta", &LatticeFasterDecoderConfig::beam_delta)
      .def_readwrite("hash_ratio", &LatticeFasterDecoderConfig::hash_ratio)
      .def_readwrite("prune_scale", &LatticeFasterDecoderConfig::prune_scale)
      .def_readwrite("memory_pool_tokens_block_size",
                     &LatticeFasterDecoderConfig::memory_pool_tokens_block_size)
      .def_readwrite("memory_pool_links_block_size",
                     &LatticeFasterDecoderConfig::memory_pool_links_block_size)
      .def("check", &LatticeFasterDecoderConfig::Check)
      .def("__str__", &LatticeFasterDecoderConfig::ToString);

  py::class_<DeviceConfig>(m, "DeviceConfig")
      .def(py::init<>())
      .def_readwrite("device", &DeviceConfig::device)
      .def_readwrite("hash_capacity", &DeviceConfig::hash_capacity)
      .def_readwrite("arena_records", &DeviceConfig::arena_records)
      .def_readwrite("threads_per_lane", &DeviceConfig::threads_per_lane)
      .def_readwrite("chunk_frames", &DeviceConfig::chunk_frames);

  py::class_<DeviceGraph, std::shared_ptr<DeviceGraph>>(m, "DeviceGraph")
      .def(py::init([](const fst::StdFst &f, int32_t device) {
             return std::make_shared<DeviceGraph>(f, device);
           }),
           py::arg("fst"), py::arg("device") = 0);

  py::class_<FasterDecoder>(m, "FasterDecoder")
      // the reference's three overloads (python/csrc/faster-decoder.cc:34-42)
      .def(py::init([](const fst::StdFst &f, const FasterDecoderOptions &config) {
             return std::make_unique<FasterDecoder>(f, config);
           }),
           py::arg("fst"), py::arg("config"))
      .def(py::init([](const fst::StdVectorFst &f, const FasterDecoderOptions &config) {
             return std::make_unique<FasterDecoder>(f, config);
           }),
           py::arg("fst"), py::arg("config"))
      .def(py::init([](const fst::StdConstFst &f, const FasterDecoderOptions &config) {
             return std::make_unique<FasterDecoder>(f, config);
           }),
           py::arg("fst"), py::arg("config"))
      .def(py::init([](std::shared_ptr<DeviceGraph> g, const FasterDecoderOptions &config,
                       const DeviceConfig &dev) {
             return std::make_unique<FasterDecoder>(std::move(g), config, dev);
           }),
           py::arg("graph"), py::arg("config"), py::arg("device_config") = DeviceConfig())
      .def("set_options", &FasterDecoder::SetOptions, py::arg("config"))
      // The GIL is released while the device works (the reference holds it for the whole
      // search): decoders driven from several Python threads -- one FasterDecoder per thread
      // on a shared DeviceGraph -- run their searches concurrently on the GPU.  A decodable
      // defined in Python re-acquires it inside its callbacks (pybind11 trampolines).
      .def("decode", &FasterDecoder::Decode, py::arg("decodable"),
           py::call_guard<py::gil_scoped_release>())
      .def("reached_final", &FasterDecoder::ReachedFinal, py::call_guard<py::gil_scoped_release>())
      .def(
          "get_best_path",
          [](FasterDecoder &self, bool use_final_probs) -> std::pair<bool, fst::Lattice> {
            fst::Lattice lat;
            bool ok;
            {
              py::gil_scoped_release nogil;
              ok = self.GetBestPath(&lat, use_final_probs);
            }
            return std::make_pair(ok, lat);
          },
          py::arg("use_final_probs") = true)
      .def("init_decoding", &FasterDecoder::InitDecoding, py::call_guard<py::gil_scoped_release>())
      .def("advance_decoding", &FasterDecoder::AdvanceDecoding, py::arg("decodable"),
           py::arg("max_num_frames") = -1, py::call_guard<py::gil_scoped_release>())
      .def("num_frames_decoded", &FasterDecoder::NumFramesDecoded);

  // ---- SimpleDecoder (kaldi-decoder/python/csrc/simple-decoder.cc:14-44)
  py::class_<SimpleDecoder>(m, "SimpleDecoder")
      .def(py::init([](const fst::StdFst &f, float beam) {
             return std::make_unique<SimpleDecoder>(f, beam);
           }),
           py::arg("fst"), py::arg("beam"))
      .def(py::init([](std::shared_ptr<DeviceGraph> g, float beam, const DeviceConfig &dev) {
             return std::make_unique<SimpleDecoder>(std::move(g), beam, dev);
           }),
           py::arg("graph"), py::arg("beam"), py::arg("device_config") = DeviceConfig())
      .def("decode", &SimpleDecoder::Decode, py::arg("decodable"))
      .def("reached_final", &SimpleDecoder::ReachedFinal)
      .def(
          "get_best_path",
          [](SimpleDecoder &self, bool use_final_probs) -> std::pair<bool, fst::Lattice> {
            fst::Lattice lat;
            bool ok = self.GetBestPath(&lat, use_final_probs);
            return std::make_pair(ok, lat);
          },
          py::arg("use_final_probs") = true)
      .def("final_relative_cost", &SimpleDecoder::FinalRelativeCost)
      .def("init_decoding", &SimpleDecoder::InitDecoding)
      .def("advance_decoding", &SimpleDecoder::AdvanceDecoding, py::arg("decodable"),
           py::arg("max_num_frames") = -1)
      .def("num_frames_decoded", &SimpleDecoder::NumFramesDecoded);

  // ---- additive: many lanes per call
  py::class_<BatchFasterDecoder>(m, "BatchFasterDecoder")
      .def(py::init([](const fst::StdFst &f, const FasterDecoderOptions &config,
                       int32_t max_lanes, const DeviceConfig &dev) {
             return std::make_unique<BatchFasterDecoder>(f, config, max_lanes, dev);
           }),
           py::arg("fst"), py::arg("config"), py::arg("max_lanes"),
           py::arg("device_config") = DeviceConfig())
      .def_property_readonly("max_lanes", &BatchFasterDecoder::MaxLanes)
      .def("set_options", &BatchFasterDecoder::SetOptions, py::arg("config"))
      .def("init_decoding", &BatchFasterDecoder::InitDecoding, py::arg("lanes"))
      .def(
          "advance_decoding",
          [](BatchFasterDecoder &self, const std::vector<int32_t> &lanes,
             const std::vector<FloatArray> &feats, const std::vector<int32_t> &offsets,
             int32_t max_num_frames) {
            if (feats.size() != lanes.size())
              throw std::runtime_error("advance_decoding: one matrix per lane expected");
            std::vector<const float *> mats;
            std::vector<int32_t> rows;
            int32_t cols = 0;
            for (const auto &a : feats) {
              if (a.ndim() != 2) throw std::runtime_error("advance_decoding: matrices must be 2-D");
              if (cols == 0) cols = static_cast<int32_t>(a.shape(1));
              if (a.shape(1) != cols) throw std::runtime_error("advance_decoding: column mismatch");
              mats.push_back(a.data());
              rows.push_back(static_cast<int32_t>(a.shape(0)));
            }
            py::gil_scoped_release nogil;
            self.AdvanceDecoding(lanes, mats, rows, cols, offsets, max_num_frames, false);
          },
          py::arg("lanes"), py::arg("feats"), py::arg("offsets") = std::vector<int32_t>(),
          py::arg("max_num_frames") = -1)
      .def(
          "advance_decoding_ptrs",
          [](BatchFasterDecoder &self, const std::vector<int32_t> &lanes,
             const std::vector<uintptr_t> &ptrs, const std::vector<int32_t> &rows, int32_t cols,
             const std::vector<int32_t> &offsets, int32_t max_num_frames, bool device_memory) {
            std::vector<const float *> mats;
            for (auto p : ptrs) mats.push_back(reinterpret_cast<const float *>(p));
            py::gil_scoped_release nogil;
            self.AdvanceDecoding(lanes, mats, rows, cols, offsets, max_num_frames, device_memory);
          },
          py::arg("lanes"), py::arg("ptrs"), py::arg("rows"), py::arg("cols"),
          py::arg("offsets") = std::vector<int32_t>(), py::arg("max_num_frames") = -1,
          py::arg("device_memory") = false,
          "Raw float32 row-major matrices by address (e.g. torch tensor.data_ptr(), host or CUDA)")
      .def(
          "decode_async",
          [](BatchFasterDecoder &self, const std::vector<int32_t> &lanes,
             const std::vector<uintptr_t> &ptrs, const std::vector<int32_t> &rows, int32_t cols,
             bool device_memory, uintptr_t producer_stream) {
            std::vector<const float *> mats;
            for (auto p : ptrs) mats.push_back(reinterpret_cast<const float *>(p));
            py::gil_scoped_release nogil;
            return self.DecodeAsync(lanes, mats, rows, cols, device_memory,
                                    reinterpret_cast<void *>(producer_stream));
          },
          py::arg("lanes"), py::arg("ptrs"), py::arg("rows"), py::arg("cols"),
          py::arg("device_memory") = false, py::arg("producer_stream") = 0,
          "Deferred InitDecoding + AdvanceDecoding over all rows + GetBestPath of the lanes, one "
          "kernel launch; returns a ticket at once.  The matrices (raw float32 row-major, by "
          "address) must stay valid until wait()/get_results().  producer_stream: the CUDA "
          "stream device matrices were produced on (e.g. torch.cuda.current_stream().cuda_stream)")
      .def(
          "wait",
          [](BatchFasterDecoder &self, int64_t ticket) {
            py::gil_scoped_release nogil;
            self.Wait(ticket);
          },
          py::arg("ticket") = -1)
      .def(
          "get_results",
          [](BatchFasterDecoder &self, int64_t ticket, bool use_final_probs) {
            std::vector<int32_t> lanes;
            std::vector<fst::Lattice> lats;
            std::vector<bool> ok;
            self.GetResults(ticket, &lanes, &lats, &ok, use_final_probs);
            return py::make_tuple(lanes, ok, lats);
          },
          py::arg("ticket"), py::arg("use_final_probs") = true,
          "(lanes, ok, lattices) of a decode_async call, without launching anything")
      .def("num_frames_decoded", &BatchFasterDecoder::NumFramesDecoded, py::arg("lane"))
      .def("reached_final", &BatchFasterDecoder::ReachedFinal, py::arg("lane"))
      .def(
          "get_best_path",
          [](BatchFasterDecoder &self, int32_t lane, bool use_final_probs) {
            fst::Lattice lat;
            bool ok = self.GetBestPath(lane, &lat, use_final_probs);
            return std::make_pair(ok, lat);
          },
          py::arg("lane"), py::arg("use_final_probs") = true)
      .def(
          "get_best_paths",
          [](BatchFasterDecoder &self, const std::vector<int32_t> &lanes, bool use_final_probs) {
            std::vector<fst::Lattice> lats;
            std::vector<bool> ok;
            self.GetBestPaths(lanes, &lats, &ok, use_final_probs);
            return std::make_pair(ok, lats);
          },
          py::arg("lanes"), py::arg("use_final_probs") = true);

  // Test hook: fst::RemoveEpsLocal (minifst restatement of kaldifst's, the last
  // step of GetBestPath, faster-decoder.cc:422) applied to a linear lattice.
  m.def(
      "_remove_eps_local_linear",
      [](const std::vector<int32_t> &il, const std::vector<int32_t> &ol,
         const std::vector<float> &gw, const std::vector<float> &aw, std::pair<float, float> fin) {
        fst::Lattice lat;
        int cur = lat.AddState();
        lat.SetStart(cur);
        for (size_t i = 0; i < il.size(); ++i) {
          int nxt = lat.AddState();
          lat.AddArc(cur, fst::LatticeArc(il[i], ol[i], fst::LatticeWeight(gw[i], aw[i]), nxt));
          cur = nxt;
        }
        lat.SetFinal(cur, fst::LatticeWeight(fin.first, fin.second));
        fst::RemoveEpsLocal(&lat);
        return lat;
      },
      py::arg("ilabels"), py::arg("olabels"), py::arg("graph"), py::arg("acoustic"),
      py::arg("final"));

  m.def("graph_uploads", &DeviceGraph::NumUploads,
        "How many graphs this process has converted and copied to a GPU so far.");
  m.def("clear_graph_cache", &DeviceGraph::ClearCache,
        "Lets go of the device graphs kept for decoders that are constructed from an FST "
        "(DeviceGraph::Shared); graphs still used by a decoder live on until it is gone.");
  m.def("device_count", []() {
    int n = 0;
    kd_device_count(&n);
    return n;
  });
}
