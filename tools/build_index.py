#!/usr/bin/env python3
"""CLI twin of the reference's scripts/build_faiss_index.py (same flags) for the B200 index.

    python tools/build_index.py --model-path M --data-path docs.parquet --output-dir artifacts/index \
        [--max-docs N] [--batch-size 32] [--device cuda] [--hnsw-m 32] [--hnsw-ef-construction 200]
    python tools/build_index.py --embeddings emb.npy [--doc-ids ids.json] --output-dir artifacts/index

--model-path is loaded with sentence-transformers if that package is present (the reference's
StudentModel wraps one: SURVEY.md App. A); "passage: " is prefixed for e5 models like the
reference's encode_documents.  --embeddings skips encoding (north star: "build from an embedding
array").  --hnsw-* are accepted and ignored: the index is exact, its build is a bf16 copy.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from semantic_search_kd_b200 import FAISSIndexBuilder  # noqa: E402


class _SentenceTransformerStudent:
    def __init__(self, path, device):
        from sentence_transformers import SentenceTransformer   # optional dependency, not in this image
        self.model = SentenceTransformer(path, device=device)
        self.e5 = "e5" in str(path).lower()
        self.embedding_dim = self.model.get_sentence_embedding_dimension()

    def encode_documents(self, texts, **kw):
        if self.e5:
            texts = ["passage: " + t for t in texts]
        return self.model.encode(texts, convert_to_numpy=True, normalize_embeddings=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", type=str)
    ap.add_argument("--data-path", type=str)
    ap.add_argument("--embeddings", type=str, help=".npy [N, dim] array to index instead of encoding a parquet")
    ap.add_argument("--doc-ids", type=str, help="JSON list of ids for --embeddings")
    ap.add_argument("--output-dir", type=str, required=True)
    ap.add_argument("--max-docs", type=int, default=None)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--device", type=str, default="cuda")
    ap.add_argument("--hnsw-m", type=int, default=32)
    ap.add_argument("--hnsw-ef-construction", type=int, default=200)
    args = ap.parse_args()
    if args.batch_size <= 0 or (args.max_docs is not None and args.max_docs <= 0):
        ap.error("--batch-size and --max-docs must be positive")
    t0 = time.time()
    if args.embeddings:
        emb = np.load(args.embeddings, mmap_mode="r")
        if args.max_docs:
            emb = emb[: args.max_docs]
        ids = json.loads(Path(args.doc_ids).read_text())[: len(emb)] if args.doc_ids else None
        builder = FAISSIndexBuilder(embedding_dim=emb.shape[1], index_type="HNSW", metric="cosine")
        for s in range(0, len(emb), 1 << 18):
            builder.add(np.asarray(emb[s:s + (1 << 18)], dtype=np.float32), ids[s:s + (1 << 18)] if ids else None)
        index = builder.index
    else:
        if not args.model_path or not args.data_path:
            ap.error("--model-path and --data-path are required unless --embeddings is given")
        for pth, flag in ((args.model_path, "--model-path"), (args.data_path, "--data-path")):
            if not Path(pth).exists():
                ap.error(f"{flag}: {pth} does not exist")
        model = _SentenceTransformerStudent(args.model_path, args.device)
        builder = FAISSIndexBuilder(embedding_dim=model.embedding_dim, index_type="HNSW", metric="cosine")
        index = builder.build_from_parquet(model=model, parquet_path=Path(args.data_path), batch_size=args.batch_size,
                                           max_docs=args.max_docs, hnsw_m=args.hnsw_m,
                                           hnsw_ef_construction=args.hnsw_ef_construction)
    builder.save(Path(args.output_dir))
    print(f"Index saved to: {args.output_dir}\nTotal vectors: {index.ntotal}\nSeconds: {time.time() - t0:.2f}")


if __name__ == "__main__":
    main()
