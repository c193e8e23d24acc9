"""README-level surface: ProgressiveSearchEngine.search and RAGSystem.search / add_document.

The reference documents these names (README.md:91-102,153-188; docs/API_GUIDE.md:55-93)
but ships no working implementation of them, so their results are "parity
unpinned"; they are thin shims that route vector work to the pinned device path
(search.search_batch).  Text -> embedding is host plumbing: pass `embed_fn`, or use the
deterministic hashing embedder below (a placeholder, not a language model).
"""
from __future__ import annotations

import hashlib
import re
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from .search import EmbeddingDatabase, search_batch


@dataclass
class DocumentSearchResult:
    """Fields the README reads from a result (README.md:98-102)."""
    document_id: str
    similarity_score: float
    content: str
    frame_number: int = -1


class HashingEmbedder:
    """Deterministic bag-of-tokens feature hashing -> unit vector (host plumbing)."""

    def __init__(self, dimension: int = 1536):
        self.dimension = int(dimension)

    def __call__(self, text: str) -> np.ndarray:
        v = np.zeros(self.dimension, dtype=np.float32)
        for tok in re.findall(r"\w+", text.lower()):
            h = int.from_bytes(hashlib.blake2b(tok.encode(), digest_size=8).digest(), "little")
            v[h % self.dimension] += 1.0 if (h >> 63) == 0 else -1.0
        nrm = float(np.linalg.norm(v))
        return v / nrm if nrm > 0 else v


class ProgressiveSearchEngine:
    """`ProgressiveSearchEngine(use_frame_caching=True[, cache_size])`, `.add_document(doc)`,
    `.search(query, max_results=10)` (docs/API_GUIDE.md:55-76).  Documents are anything with
    an `.embedding` (or a dict with "embedding"), or raw vectors; queries are vectors, batches
    of vectors, or text when `embed_fn` is given."""

    def __init__(self, use_frame_caching: bool = True, cache_size: int = 1000, embed_fn: Optional[Callable] = None,
                 device=None, use_filter: bool = True):
        self.use_frame_caching = use_frame_caching
        self.cache_size = cache_size
        self.embed_fn = embed_fn
        self._device = device
        self.use_filter = use_filter
        self._vectors: List[np.ndarray] = []
        self._docs: list = []
        self._db: Optional[EmbeddingDatabase] = None
        self._bulk: Optional[torch.Tensor] = None

    def _vector_of(self, doc) -> np.ndarray:
        if isinstance(doc, str):
            if self.embed_fn is None:
                raise ValueError("text needs an embed_fn")
            return np.asarray(self.embed_fn(doc), dtype=np.float32)
        if isinstance(doc, dict) and "embedding" in doc:
            return np.asarray(doc["embedding"], dtype=np.float32)
        if hasattr(doc, "embedding"):
            return np.asarray(doc.embedding, dtype=np.float32)
        return np.asarray(doc, dtype=np.float32)

    def add_document(self, doc) -> int:
        self._vectors.append(self._vector_of(doc).reshape(-1))
        self._docs.append(doc)
        self._db = None
        return len(self._docs) - 1

    def add_embeddings(self, embeddings) -> None:
        """Bulk load [N, D] (ndarray or device tensor)."""
        if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda and not self._vectors:
            self._db = EmbeddingDatabase(embeddings, device=embeddings.device)
            self._docs = list(range(self._db.N))
            self._bulk = embeddings                     # device rows of the bulk load: a later add_document rebuilds from them
            self._vectors = [None] * self._db.N
            return
        arr = embeddings.detach().cpu().numpy() if isinstance(embeddings, torch.Tensor) else np.asarray(embeddings)
        for row in arr:
            self.add_document(row)

    def _database(self) -> EmbeddingDatabase:
        if self._db is None:
            if not self._vectors:
                raise ValueError("no documents")
            n_bulk = 0 if self._bulk is None else int(self._bulk.shape[0])
            tail = self._vectors[n_bulk:]
            if n_bulk:                                  # bulk-loaded device rows first, documents added since then after them
                rows = self._bulk.to(torch.float32)
                if tail:
                    rows = torch.cat([rows, torch.from_numpy(np.stack(tail).astype(np.float32)).to(rows.device)])
                self._db = EmbeddingDatabase(rows, device=rows.device)
            else:
                self._db = EmbeddingDatabase(np.stack(tail), device=self._device)
        return self._db

    def search_vectors(self, queries, max_results: int = 10):
        """Batched device path: ([Q, k] ids, [Q, k] scores) tensors."""
        return search_batch(self._database(), queries, max_results, use_filter=self.use_filter)

    def search(self, query, max_results: int = 10) -> List[DocumentSearchResult]:
        q = self._vector_of(query)
        ids, scores = self.search_vectors(q.reshape(1, -1), max_results)
        out = []
        for i, s in zip(ids[0].cpu().tolist(), scores[0].cpu().tolist()):
            if i < 0:
                continue
            doc = self._docs[i]
            doc_id = getattr(doc, "document_id", None) or (doc.get("document_id") if isinstance(doc, dict) else None) or str(i)
            content = getattr(doc, "content", None) or (doc.get("content") if isinstance(doc, dict) else None) or ""
            out.append(DocumentSearchResult(str(doc_id), float(s), str(content), int(i)))
        return out


class RAGSystem:
    """`RAGSystem(config)`, `.add_document(document_id, content)`, `.search(query, max_results=10)`
    (docs/API_GUIDE.md:86-93, README.md:91-102)."""

    def __init__(self, config=None, embed_fn: Optional[Callable[[str], Sequence[float]]] = None, device=None,
                 embedding_dimension: Optional[int] = None):
        self.config = config
        dim = embedding_dimension or getattr(config, "embedding_dimension", None) or 1536
        self.embed_fn = embed_fn or HashingEmbedder(dim)
        self.engine = ProgressiveSearchEngine(embed_fn=self.embed_fn, device=device)

    def add_document(self, document_id: str, content: str) -> int:
        return self.engine.add_document({"document_id": document_id, "content": content,
                                         "embedding": np.asarray(self.embed_fn(content), dtype=np.float32)})

    def search(self, query: Union[str, np.ndarray], max_results: int = 10) -> List[DocumentSearchResult]:
        return self.engine.search(query, max_results)

    # names the real reference class uses (rag/api.py:127,268)
    def search_similar_documents(self, query_text: str, max_results: int = 10) -> List[DocumentSearchResult]:
        return self.search(query_text, max_results)
