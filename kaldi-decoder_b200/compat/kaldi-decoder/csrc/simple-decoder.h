// Forwarding header: lets a caller keep `#include "kaldi-decoder/csrc/simple-decoder.h"`.
#include "kaldi-decoder_b200/csrc/simple-decoder.h"
