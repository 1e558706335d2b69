"""Holds the compiled extension module _kaldi_decoder (built by kaldi-decoder_b200/build.py)."""
