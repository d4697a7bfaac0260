#!/usr/bin/env python
"""Tokenise a text corpus into the flat token file the pre-training scripts read with ``--data_path`` (role of the reference's
``examples/training/llama/get_dataset.py``, which downloads wikicorpus and tokenises / chunks it with ``datasets``; there is no
network here, so the corpus is local files).

    python get_dataset.py --input corpus.txt more.jsonl --tokenizer /path/to/hf_tokenizer_dir --output tokens.bin
    python get_dataset.py --input corpus.txt --tokenizer bytes --output tokens.bin          # no vocabulary file needed

Inputs: plain text (one document per blank-line separated paragraph, or ``--one_doc_per_line``) and ``.jsonl`` (field
``--text_key``).  Every document is followed by the EOS id; the stream is written as ``uint16`` (``uint32`` when the vocabulary does
not fit) next to ``<output>.json`` (vocab size, dtype, token count, tokenizer).  ``memmap_batches`` (``training_utils.py``) cuts it
into ``[batch, seq_len]`` windows strided over the data-parallel ranks."""
import argparse
import json
import os
from typing import Iterator, List

import numpy as np


class _Bytes:
    vocab_size, eos_id = 258, 257

    def encode(self, text: str) -> List[int]:
        return list(text.encode("utf-8"))


class _SentencePiece:
    def __init__(self, path: str):
        import sentencepiece as spm

        self.sp = spm.SentencePieceProcessor(model_file=path)
        self.vocab_size, self.eos_id = self.sp.vocab_size(), self.sp.eos_id()

    def encode(self, text: str) -> List[int]:
        return self.sp.encode(text)


class _HF:
    def __init__(self, path: str):
        from transformers import AutoTokenizer

        self.tok = AutoTokenizer.from_pretrained(path, local_files_only=True)
        self.vocab_size = len(self.tok)
        self.eos_id = self.tok.eos_token_id if self.tok.eos_token_id is not None else 0

    def encode(self, text: str) -> List[int]:
        return self.tok.encode(text, add_special_tokens=False)


def load_tokenizer(spec: str):
    if spec == "bytes":
        return _Bytes()
    if spec.endswith(".model"):
        return _SentencePiece(spec)
    return _HF(spec)


def documents(paths: List[str], text_key: str, one_doc_per_line: bool) -> Iterator[str]:
    for path in paths:
        with open(path, encoding="utf-8") as f:
            if path.endswith(".jsonl"):
                for line in f:
                    if line.strip():
                        yield json.loads(line)[text_key]
            elif one_doc_per_line:
                for line in f:
                    if line.strip():
                        yield line.rstrip("\n")
            else:
                para: List[str] = []
                for line in f:
                    if line.strip():
                        para.append(line.rstrip("\n"))
                    elif para:
                        yield "\n".join(para)
                        para = []
                if para:
                    yield "\n".join(para)


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--input", nargs="+", required=True)
    p.add_argument("--tokenizer", default="bytes", help='"bytes", a SentencePiece .model file, or a local HF tokenizer directory')
    p.add_argument("--output", required=True)
    p.add_argument("--text_key", default="text")
    p.add_argument("--one_doc_per_line", action="store_true")
    p.add_argument("--min_tokens", type=int, default=0, help="fail if the corpus is shorter (e.g. seq_len * batch)")
    a = p.parse_args(argv)
    tok = load_tokenizer(a.tokenizer)
    dtype = np.uint16 if tok.vocab_size <= 65536 else np.uint32
    chunks, n_docs, n_tok = [], 0, 0
    for doc in documents(a.input, a.text_key, a.one_doc_per_line):
        ids = tok.encode(doc) + [tok.eos_id]
        chunks.append(np.asarray(ids, dtype=dtype))
        n_docs, n_tok = n_docs + 1, n_tok + len(ids)
    if n_tok < max(2, a.min_tokens):
        raise SystemExit(f"corpus has {n_tok} tokens, need at least {max(2, a.min_tokens)}")
    os.makedirs(os.path.dirname(os.path.abspath(a.output)), exist_ok=True)
    np.concatenate(chunks).tofile(a.output)
    meta = {"tokens": n_tok, "documents": n_docs, "dtype": np.dtype(dtype).name, "vocab_size": tok.vocab_size, "eos_id": tok.eos_id,
            "tokenizer": a.tokenizer}
    with open(a.output + ".json", "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta))
    return meta


if __name__ == "__main__":
    main()
