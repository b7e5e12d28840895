"""Fold assignment for k-fold training / fold-sharded ensemble evaluation -- the product-side counterpart of
`generate_kfold_splits` (scripts/prepare_kfold_data.py:30-73).

The reference draws its folds with scikit-learn's `StratifiedKFold(k, shuffle=True, random_state=42)` and the index
lists it committed (`data/splits/split_fold_{1..7}.json`) must be reproduced bit for bit, so the same scikit-learn
call is the fold source here (it is a pinned dependency of the reference, requirements.txt:144); everything around
it -- the rotating test / val / train rule, the JSON layout and file names the reference's DataModule reads, the
class-ordered label vector of the CARS dataset -- is this module's.  Host-side integer work only: no GPU involved.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

__all__ = ["stratified_folds", "rotate_folds", "generate_kfold_splits", "cars_labels", "load_fold_split"]


def cars_labels(n_normal: int = 225, n_cancerous: int = 225) -> np.ndarray:
    """Label vector of the CARS thyroid set in the order `CARSThyroidDataset(split='all')` enumerates it: every
    'normal' image (label 0) first, then every 'cancerous' one (label 1) -- src/data/dataset.py:96."""
    return np.concatenate([np.zeros(n_normal, dtype=np.int64), np.ones(n_cancerous, dtype=np.int64)])


def stratified_folds(labels: Sequence[int], k: int, random_state: int = 42) -> List[np.ndarray]:
    """The k disjoint held-out index sets (:48-49): fold j = the test indices of the j-th StratifiedKFold split."""
    from sklearn.model_selection import StratifiedKFold
    labels = np.asarray(labels)
    if labels.ndim != 1:
        raise ValueError("labels must be a 1-D sequence")
    if k < 3:
        raise ValueError("rotating train/val/test folds need k >= 3")     # test, val and at least one training fold
    skf = StratifiedKFold(n_splits=k, shuffle=True, random_state=random_state)
    return [held_out for _, held_out in skf.split(np.arange(len(labels)), labels)]


def rotate_folds(folds: Sequence[np.ndarray]) -> List[Dict[str, List[int]]]:
    """Run i tests on fold i, validates on fold (i+1) mod k and trains on the rest, concatenated in fold order (:52-63)."""
    k = len(folds)
    runs = []
    for i in range(k):
        v = (i + 1) % k
        train = np.concatenate([folds[j] for j in range(k) if j != i and j != v])
        runs.append({"train": [int(x) for x in train], "val": [int(x) for x in folds[v]], "test": [int(x) for x in folds[i]]})
    return runs


def generate_kfold_splits(labels_or_data_dir: Union[Sequence[int], str, Path], k: int, random_state: int = 42,
                          splits_dir: Optional[Union[str, Path]] = None) -> List[Dict[str, List[int]]]:
    """scripts/prepare_kfold_data.py:30-73.

    `labels_or_data_dir`: the label vector, or (the reference's calling convention) the raw-data directory holding
    `normal/` and `cancerous/` sub-directories, whose image counts give the class-ordered label vector.
    Writes `split_fold_{i}.json` ({'train','val','test'}, indent 2, 1-based file names, :65-71) into `splits_dir`
    (default for a data directory: `<data_dir>/../splits`, :38-39) and returns the k runs."""
    if isinstance(labels_or_data_dir, (str, Path)):
        data_dir = Path(labels_or_data_dir)
        counts = []
        for cls in ("normal", "cancerous"):
            d = data_dir / cls
            if not d.is_dir():
                raise FileNotFoundError(f"expected a class directory {d}")
            counts.append(sum(1 for p in d.iterdir() if p.is_file()))
        labels = cars_labels(*counts)
        if splits_dir is None:
            splits_dir = data_dir.parent / "splits"
    else:
        labels = np.asarray(labels_or_data_dir)
    runs = rotate_folds(stratified_folds(labels, k, random_state))
    if splits_dir is not None:
        out = Path(splits_dir)
        out.mkdir(parents=True, exist_ok=True)
        for i, run in enumerate(runs, start=1):
            with open(out / f"split_fold_{i}.json", "w") as f:
                json.dump(run, f, indent=2)
    return runs


def load_fold_split(splits_dir: Union[str, Path], fold: int) -> Dict[str, List[int]]:
    """Reads `split_fold_{fold}.json` (1-based, as the reference's DataModule does for `fold` runs)."""
    with open(Path(splits_dir) / f"split_fold_{fold}.json") as f:
        d = json.load(f)
    return {k: [int(x) for x in d[k]] for k in ("train", "val", "test")}
