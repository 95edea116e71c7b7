"""Mirror of the per-category alpha table and centroid routing
(src/search/router.rs:126-175, :708-833, :1326-1490).  The alpha resolution is
host logic (a table and env lookups, exactly as in the reference); the centroid
dot products run on the GPU (``cqs_b200_route_centroids``)."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Mapping, Optional

import numpy as np

from .capi import lib, check, ptr

# declaration order of define_query_categories! (router.rs:126-175)
CATEGORIES = (
    "identifier_lookup", "structural", "behavioral", "conceptual", "multi_step",
    "negation", "type_filtered", "cross_language", "unknown",
)
DEFAULT_ALPHA = {
    "identifier_lookup": 0.85, "structural": 0.60, "behavioral": 1.00, "conceptual": 0.80,
    "multi_step": 0.10, "negation": 0.80, "type_filtered": 0.00, "cross_language": 0.70,
    "unknown": 0.80,
}
ALIASES = {"structural_search": "structural", "behavioral_search": "behavioral",
           "conceptual_search": "conceptual"}
CENTROID_ALPHA_FLOOR = 0.7  # src/cli/commands/search/query.rs:655-657


def _parse_f32(val: str):
    if val != val.strip() or "_" in val or val == "":
        return None
    try:
        with np.errstate(all="ignore"):
            return np.float32(float(val))
    except ValueError:
        return None


def _clamp01(a) -> np.float32:
    return np.float32(min(max(np.float32(a), np.float32(0.0)), np.float32(1.0)))


def resolve_splade_alpha(category: str, env: Optional[Mapping[str, str]] = None,
                         slot_table: Optional[Mapping[str, float]] = None) -> np.float32:
    """per-category env > global env > slot.toml > default (router.rs:708-833)."""
    env = os.environ if env is None else env
    for key in (f"CQS_SPLADE_ALPHA_{category.upper()}", "CQS_SPLADE_ALPHA"):
        val = env.get(key)
        if val is not None:
            a = _parse_f32(val)
            if a is not None and np.isfinite(a):
                return _clamp01(a)
    if slot_table:
        a = slot_table.get(category.lower())
        if a is not None and np.isfinite(np.float32(a)):
            return _clamp01(a)
    return np.float32(DEFAULT_ALPHA[category])


def apply_centroid_floor(alpha, centroid_applied: bool) -> np.float32:
    a = np.float32(alpha)
    return np.float32(max(a, np.float32(CENTROID_ALPHA_FLOOR))) if centroid_applied else a


class CentroidClassifier:
    """router.rs:1326-1445.  ``centroids``: {category: f32[dim]}."""

    def __init__(self, centroids: Mapping[str, np.ndarray], threshold: float = 0.01, device: int = 0):
        self.names = [c for c in CATEGORIES if c in centroids]
        self.matrix = np.ascontiguousarray(np.stack([np.asarray(centroids[c], np.float32) for c in self.names]))
        self.dim = self.matrix.shape[1]
        self.threshold = float(threshold)
        self.device = device

    @classmethod
    def load(cls, path: str, env: Optional[Mapping[str, str]] = None, device: int = 0):
        """classifier_centroids.v1.json loader (router.rs:1333-1413); None = rule-only mode."""
        env = os.environ if env is None else env
        if env.get("CQS_CENTROID_CLASSIFIER") == "0":
            return None
        try:
            if os.path.getsize(path) > 16 * 1024 * 1024:
                return None
            with open(path) as f:
                data = json.load(f)
            dim = int(data["dim"])
            cents = {}
            for name, obj in data["categories"].items():
                name = ALIASES.get(name, name)
                if name not in DEFAULT_ALPHA:
                    return None
                vec = np.asarray(obj["centroid"], dtype=np.float32)
                if vec.shape[0] != dim:
                    continue
                cents[name] = vec
        except (OSError, ValueError, KeyError, TypeError):
            return None
        if not cents:
            return None
        thr = _parse_f32(env.get("CQS_CENTROID_THRESHOLD", "0.01"))
        return cls(cents, 0.01 if thr is None else float(thr), device)

    def classify_batch(self, embeddings: np.ndarray):
        """-> (categories list[str|None], margins f32[nq])"""
        q = np.ascontiguousarray(embeddings, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        if q.shape[1] != self.dim:
            return [None] * nq, np.zeros(nq, np.float32)
        cat = np.empty(nq, np.int32)
        margin = np.empty(nq, np.float32)
        check(lib.cqs_b200_route_centroids(self.device, ptr(self.matrix), self.matrix.shape[0], self.dim,
                                           ptr(q), nq, self.threshold, ptr(cat), ptr(margin)))
        return [self.names[c] if c >= 0 else None for c in cat.tolist()], margin

    def classify(self, embedding: np.ndarray):
        cats, m = self.classify_batch(embedding)
        return (cats[0], float(m[0])) if cats[0] is not None else None


def reclassify_with_centroid(category: str, embedding: np.ndarray, classifier: Optional[CentroidClassifier],
                             env: Optional[Mapping[str, str]] = None):
    """router.rs:1453-1490: only fills `unknown`.  -> (category, centroid_applied)"""
    env = os.environ if env is None else env
    if env.get("CQS_CENTROID_CLASSIFIER") == "0" or category != "unknown" or classifier is None:
        return category, False
    got = classifier.classify(embedding)
    if got is None:
        return category, False
    return got[0], True
