"""Drop-in mirror of the reference's ``rnnt`` package for the hot path only:
``rnnt.models`` (Transducer / Encoder / Decoder / Joint / TimeReduction / ResLayerNormLSTM),
``rnnt.stream`` (PytorchStreamDecoder) and the ``rnnt.tokenizer`` constants."""
