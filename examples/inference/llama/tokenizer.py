"""Tokenizers for the sample.  There is no network in the build image, so the default is a byte-level tokenizer that needs no
vocabulary file; a SentencePiece ``.model`` file is used when one is given (Llama-2 style checkpoints)."""
from typing import List, Sequence


class ByteTokenizer:
    """256 byte values + BOS / EOS / PAD / EOT.  ``vocab_size`` = 260."""

    def __init__(self):
        self.bos_id, self.eos_id, self.pad_id, self.eot_id = 256, 257, 258, 259
        self.vocab_size = 260
        self.stop_tokens = [self.eos_id, self.eot_id]

    def encode(self, text: str, bos: bool = True, eos: bool = False) -> List[int]:
        ids = list(text.encode("utf-8"))
        return ([self.bos_id] if bos else []) + ids + ([self.eos_id] if eos else [])

    def decode(self, ids: Sequence[int]) -> str:
        return bytes(i for i in ids if i < 256).decode("utf-8", errors="replace")


class SentencePieceTokenizer:
    def __init__(self, model_path: str):
        import sentencepiece as spm

        self.sp = spm.SentencePieceProcessor(model_file=model_path)
        self.bos_id, self.eos_id = self.sp.bos_id(), self.sp.eos_id()
        self.pad_id = self.sp.pad_id() if self.sp.pad_id() >= 0 else 0
        self.vocab_size = self.sp.vocab_size()
        self.stop_tokens = [self.eos_id]

    def encode(self, text: str, bos: bool = True, eos: bool = False) -> List[int]:
        ids = self.sp.encode(text)
        return ([self.bos_id] if bos else []) + ids + ([self.eos_id] if eos else [])

    def decode(self, ids: Sequence[int]) -> str:
        return self.sp.decode([i for i in ids if i != self.pad_id])


def load_tokenizer(path=None):
    return SentencePieceTokenizer(path) if path else ByteTokenizer()
