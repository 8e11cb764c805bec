"""Special token ids the model path depends on (reference: rnnt/tokenizer.py:7-10).  The BPE
tokenizer itself is host-side text processing outside the hot path and is not rebuilt here."""
NUL = 0
PAD = 1
BOS = 2
UNK = 3
