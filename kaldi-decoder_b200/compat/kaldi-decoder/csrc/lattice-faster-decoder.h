// Forwarding header: lets code written against the reference include path
// ("kaldi-decoder/csrc/lattice-faster-decoder.h") build against the B200 implementation with
// -I kaldi-decoder_b200/compat -I <repo root>.
#include "kaldi-decoder_b200/csrc/lattice-faster-decoder.h"
