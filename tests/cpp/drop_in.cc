// tests/cpp/drop_in.cc -- a C++ caller written against the REFERENCE's include paths and API
// (kaldi-decoder/csrc/faster-decoder.h:65-107, decodable-ctc.h:18-24), built with
//   -I kaldi-decoder_b200/compat -I <repo> -I kaldi-decoder_b200/csrc/minifst
// and linked against the B200 implementation.  It reads a graph and log-prob matrices written
// by the test (tests/test_gpu_api.py), decodes every utterance the way sherpa-style callers do,
// and prints one line per utterance: ok, reached_final, the output labels, the total cost.
//
//   drop_in <graph.fst> <beam> <max_active> <min_active> <logp.bin> <n_utts> <T> <V>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "kaldi-decoder/csrc/decodable-ctc.h"
#include "kaldi-decoder/csrc/faster-decoder.h"
#include "kaldi-decoder_b200/csrc/fst-io.h"  // stands in for fst::StdVectorFst::Read (no OpenFst here)

int main(int argc, char **argv) {
  if (argc != 9) {
    std::fprintf(stderr, "usage: drop_in graph.fst beam max_active min_active logp.bin n T V\n");
    return 2;
  }
  const fst::StdVectorFst graph = kaldi_decoder::ReadFst(argv[1]);
  kaldi_decoder::FasterDecoderOptions opts;  // reference defaults
  opts.beam = static_cast<float>(std::atof(argv[2]));
  opts.max_active = std::atoi(argv[3]);
  opts.min_active = std::atoi(argv[4]);
  const int n = std::atoi(argv[6]), T = std::atoi(argv[7]), V = std::atoi(argv[8]);
  std::vector<float> logp(static_cast<size_t>(n) * T * V);
  std::ifstream is(argv[5], std::ios::binary);
  is.read(reinterpret_cast<char *>(logp.data()), logp.size() * sizeof(float));
  if (!is) {
    std::fprintf(stderr, "short read of %s\n", argv[5]);
    return 2;
  }
  try {
    kaldi_decoder::FasterDecoder decoder(graph, opts);
    for (int u = 0; u < n; ++u) {
      // zero-copy decodable over the caller's matrix (decodable-ctc.h:18-24)
      kaldi_decoder::DecodableCtc decodable(logp.data() + static_cast<size_t>(u) * T * V, T, V);
      if (u % 2 == 0) {
        decoder.Decode(&decodable);
      } else {  // the streaming calls, in two pieces
        decoder.InitDecoding();
        decoder.AdvanceDecoding(&decodable, T / 3);
        decoder.AdvanceDecoding(&decodable);
      }
      fst::VectorFst<fst::LatticeArc> best;
      const bool ok = decoder.GetBestPath(&best, /*use_final_probs=*/true);
      std::vector<int32_t> isyms, osyms;
      fst::LatticeWeight w;
      kaldi_decoder::GetLinearSymbolSequence(best, &isyms, &osyms, &w);
      std::printf("%d %d %d %d %.6f", u, ok ? 1 : 0, decoder.ReachedFinal() ? 1 : 0,
                  decoder.NumFramesDecoded(), static_cast<double>(w.Value1()) + w.Value2());
      for (int32_t o : osyms) std::printf(" %d", o);
      std::printf("\n");
    }
  } catch (const std::exception &e) {  // KALDI_DECODER_ERR / ASSERT throw std::runtime_error
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
