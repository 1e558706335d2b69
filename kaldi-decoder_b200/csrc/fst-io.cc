// kaldi-decoder_b200/csrc/fst-io.cc
#include "kaldi-decoder_b200/csrc/fst-io.h"

#include <fstream>
#include <limits>
#include <sstream>

#include "kaldi-decoder_b200/csrc/log.h"

namespace kaldi_decoder {

namespace {

constexpr int32_t kFstMagic = 2125659606;
constexpr int32_t kFlagHasIsyms = 1, kFlagHasOsyms = 2, kFlagIsAligned = 4;

template <class T>
T ReadPod(std::istream &is) {
  T v;
  is.read(reinterpret_cast<char *>(&v), sizeof(T));
  if (!is) KALDI_DECODER_ERR << "unexpected end of FST file";
  return v;
}

template <class T>
void WritePod(std::ostream &os, const T &v) {
  os.write(reinterpret_cast<const char *>(&v), sizeof(T));
}

std::string ReadString(std::istream &is) {
  int32_t n = ReadPod<int32_t>(is);
  if (n < 0 || n > (1 << 20)) KALDI_DECODER_ERR << "bad string length in FST header";
  std::string s(static_cast<size_t>(n), '\0');
  if (n) is.read(&s[0], n);
  if (!is) KALDI_DECODER_ERR << "unexpected end of FST file";
  return s;
}

void WriteString(std::ostream &os, const std::string &s) {
  WritePod<int32_t>(os, static_cast<int32_t>(s.size()));
  os.write(s.data(), static_cast<std::streamsize>(s.size()));
}

// OpenFst SymbolTable binary form (symbol-table.cc, SymbolTableImpl::Read): magic int32,
// name string, available_key int64, size int64, then size x {symbol string, key int64}.
constexpr int32_t kSymbolTableMagic = 2125658996;

void SkipSymbolTable(std::istream &is) {
  if (ReadPod<int32_t>(is) != kSymbolTableMagic)
    KALDI_DECODER_ERR << "bad symbol table in FST file";
  (void)ReadString(is);          // name
  (void)ReadPod<int64_t>(is);    // available_key
  const int64_t size = ReadPod<int64_t>(is);
  if (size < 0) KALDI_DECODER_ERR << "bad symbol table size in FST file";
  for (int64_t i = 0; i < size; ++i) {
    (void)ReadString(is);
    (void)ReadPod<int64_t>(is);
  }
}

void AlignInput(std::istream &is, int64_t align = 16) {
  int64_t pos = static_cast<int64_t>(is.tellg());
  if (pos < 0) return;
  int64_t pad = (align - pos % align) % align;
  is.ignore(pad);
}

}  // namespace

fst::StdVectorFst ReadFstBinary(std::istream &is) {
  if (ReadPod<int32_t>(is) != kFstMagic) KALDI_DECODER_ERR << "not an OpenFst binary file";
  const std::string fst_type = ReadString(is);
  const std::string arc_type = ReadString(is);
  const int32_t version = ReadPod<int32_t>(is);
  const int32_t flags = ReadPod<int32_t>(is);
  (void)ReadPod<uint64_t>(is);  // properties
  const int64_t start = ReadPod<int64_t>(is);
  const int64_t num_states = ReadPod<int64_t>(is);
  const int64_t num_arcs = ReadPod<int64_t>(is);
  if (arc_type != "standard") KALDI_DECODER_ERR << "unsupported arc type: " << arc_type;
  // Embedded symbol tables follow the header (OpenFst FstImpl::ReadHeader); the decoder works
  // on integer labels, so they are skipped.
  if (flags & kFlagHasIsyms) SkipSymbolTable(is);
  if (flags & kFlagHasOsyms) SkipSymbolTable(is);
  fst::StdVectorFst out;
  if (fst_type == "vector") {
    out.ReserveStates(static_cast<size_t>(std::max<int64_t>(num_states, 0)));
    // num_states may be unset (-1) in streamed files: then read until EOF
    for (int64_t s = 0; num_states < 0 || s < num_states; ++s) {
      float fin;
      is.read(reinterpret_cast<char *>(&fin), sizeof(fin));
      if (!is) {
        if (num_states < 0) break;
        KALDI_DECODER_ERR << "unexpected end of FST file";
      }
      const int64_t narcs = ReadPod<int64_t>(is);
      const int st = out.AddState();
      out.SetFinal(st, fst::TropicalWeight(fin));
      out.ReserveArcs(st, static_cast<size_t>(narcs));
      for (int64_t a = 0; a < narcs; ++a) {
        const int32_t il = ReadPod<int32_t>(is), ol = ReadPod<int32_t>(is);
        const float w = ReadPod<float>(is);
        const int32_t ns = ReadPod<int32_t>(is);
        out.AddArc(st, fst::StdArc(il, ol, fst::TropicalWeight(w), ns));
      }
    }
  } else if (fst_type == "const") {
    const bool aligned = (flags & kFlagIsAligned) != 0 || version == 1;
    if (aligned) AlignInput(is);
    struct ConstState {
      float final;
      uint32_t pos, narcs, niepsilons, noepsilons;
    };
    std::vector<ConstState> states(static_cast<size_t>(num_states));
    for (auto &st : states) {
      st.final = ReadPod<float>(is);
      st.pos = ReadPod<uint32_t>(is);
      st.narcs = ReadPod<uint32_t>(is);
      st.niepsilons = ReadPod<uint32_t>(is);
      st.noepsilons = ReadPod<uint32_t>(is);
    }
    if (aligned) AlignInput(is);
    std::vector<fst::StdArc> arcs;
    arcs.reserve(static_cast<size_t>(num_arcs));
    for (int64_t a = 0; a < num_arcs; ++a) {
      const int32_t il = ReadPod<int32_t>(is), ol = ReadPod<int32_t>(is);
      const float w = ReadPod<float>(is);
      const int32_t ns = ReadPod<int32_t>(is);
      arcs.emplace_back(il, ol, fst::TropicalWeight(w), ns);
    }
    for (int64_t s = 0; s < num_states; ++s) {
      const int st = out.AddState();
      out.SetFinal(st, fst::TropicalWeight(states[s].final));
      for (uint32_t a = 0; a < states[s].narcs; ++a) out.AddArc(st, arcs[states[s].pos + a]);
    }
  } else {
    KALDI_DECODER_ERR << "unsupported FST type: " << fst_type;
  }
  if (start >= 0) out.SetStart(static_cast<int>(start));
  return out;
}

fst::StdVectorFst ReadFst(const std::string &path) {
  std::ifstream is(path, std::ios::binary);
  if (!is) KALDI_DECODER_ERR << "cannot open " << path;
  return ReadFstBinary(is);
}

void WriteFstBinary(const fst::Fst<fst::StdArc> &fst, std::ostream &os) {
  const int32_t n = fst::CountStates(fst);
  int64_t num_arcs = 0;
  for (int32_t s = 0; s < n; ++s) num_arcs += static_cast<int64_t>(fst.NumArcs(s));
  WritePod<int32_t>(os, kFstMagic);
  WriteString(os, "vector");
  WriteString(os, "standard");
  WritePod<int32_t>(os, 2);           // version
  WritePod<int32_t>(os, 0);           // flags: no symbol tables, not aligned
  WritePod<uint64_t>(os, 0x3ull);     // properties: expanded | mutable
  WritePod<int64_t>(os, fst.Start());
  WritePod<int64_t>(os, n);
  WritePod<int64_t>(os, num_arcs);
  for (int32_t s = 0; s < n; ++s) {
    WritePod<float>(os, fst.Final(s).Value());
    WritePod<int64_t>(os, static_cast<int64_t>(fst.NumArcs(s)));
    for (fst::ArcIterator<fst::Fst<fst::StdArc>> aiter(fst, s); !aiter.Done(); aiter.Next()) {
      const fst::StdArc &arc = aiter.Value();
      WritePod<int32_t>(os, arc.ilabel);
      WritePod<int32_t>(os, arc.olabel);
      WritePod<float>(os, arc.weight.Value());
      WritePod<int32_t>(os, arc.nextstate);
    }
  }
}

void WriteFst(const fst::Fst<fst::StdArc> &fst, const std::string &path) {
  std::ofstream os(path, std::ios::binary);
  if (!os) KALDI_DECODER_ERR << "cannot open " << path << " for writing";
  WriteFstBinary(fst, os);
}

fst::StdVectorFst ReadFstText(const std::string &text, bool acceptor) {
  fst::StdVectorFst out;
  auto ensure = [&out](int s) {
    while (out.NumStates() <= s) out.AddState();
  };
  std::istringstream lines(text);
  std::string line;
  bool have_start = false;
  while (std::getline(lines, line)) {
    std::istringstream ls(line);
    std::vector<std::string> f;
    std::string tok;
    while (ls >> tok) f.push_back(tok);
    if (f.empty()) continue;
    const int src = std::stoi(f[0]);
    ensure(src);
    if (!have_start) {
      out.SetStart(src);
      have_start = true;
    }
    auto weight_of = [](const std::string &s) -> float {
      if (s == "Infinity" || s == "inf") return std::numeric_limits<float>::infinity();
      return std::stof(s);
    };
    if (f.size() <= 2) {  // final state
      out.SetFinal(src, fst::TropicalWeight(f.size() == 2 ? weight_of(f[1]) : 0.0f));
      continue;
    }
    const int dst = std::stoi(f[1]);
    ensure(dst);
    int il, ol;
    size_t wpos;
    if (acceptor) {
      il = ol = std::stoi(f[2]);
      wpos = 3;
    } else {
      if (f.size() < 4) KALDI_DECODER_ERR << "bad FST text line: " << line;
      il = std::stoi(f[2]);
      ol = std::stoi(f[3]);
      wpos = 4;
    }
    const float w = f.size() > wpos ? weight_of(f[wpos]) : 0.0f;
    out.AddArc(src, fst::StdArc(il, ol, fst::TropicalWeight(w), dst));
  }
  return out;
}

std::string WriteFstText(const fst::Fst<fst::StdArc> &fst) {
  std::ostringstream os;
  os.precision(9);
  const int32_t n = fst::CountStates(fst);
  auto dump = [&](int32_t s) {
    for (fst::ArcIterator<fst::Fst<fst::StdArc>> aiter(fst, s); !aiter.Done(); aiter.Next()) {
      const fst::StdArc &arc = aiter.Value();
      os << s << " " << arc.nextstate << " " << arc.ilabel << " " << arc.olabel;
      if (arc.weight.Value() != 0.0f) os << " " << arc.weight.Value();
      os << "\n";
    }
    if (fst.Final(s) != fst::TropicalWeight::Zero()) {
      os << s;
      if (fst.Final(s).Value() != 0.0f) os << " " << fst.Final(s).Value();
      os << "\n";
    }
  };
  const int32_t start = fst.Start();
  if (start >= 0) dump(start);  // the start state's lines come first
  for (int32_t s = 0; s < n; ++s)
    if (s != start) dump(s);
  return os.str();
}

bool GetLinearSymbolSequence(const fst::Fst<fst::LatticeArc> &fst, std::vector<int32_t> *isyms,
                             std::vector<int32_t> *osyms, fst::LatticeWeight *total) {
  if (isyms) isyms->clear();
  if (osyms) osyms->clear();
  fst::LatticeWeight tot = fst::LatticeWeight::One();
  int s = fst.Start();
  if (s == fst::kNoStateId) {
    if (total) *total = fst::LatticeWeight::Zero();
    return false;
  }
  const int n = fst::CountStates(fst);
  for (int steps = 0; steps <= n; ++steps) {
    const fst::LatticeWeight fin = fst.Final(s);
    const size_t narcs = fst.NumArcs(s);
    if (fin != fst::LatticeWeight::Zero()) {
      if (narcs != 0) return false;  // final state with arcs: not linear
      tot = fst::Times(tot, fin);
      if (total) *total = tot;
      return true;
    }
    if (narcs != 1) return false;
    fst::ArcIterator<fst::Fst<fst::LatticeArc>> aiter(fst, s);
    const fst::LatticeArc &a = aiter.Value();
    tot = fst::Times(tot, a.weight);
    if (a.ilabel != 0 && isyms) isyms->push_back(a.ilabel);
    if (a.olabel != 0 && osyms) osyms->push_back(a.olabel);
    s = a.nextstate;
  }
  return false;  // cycle
}

}  // namespace kaldi_decoder
