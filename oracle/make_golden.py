"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in the build container.

    python -m oracle.make_golden            # needs /root/reference; rewrites tests/golden

Each fixture stores the inputs (so nothing depends on RNG stream stability) and the outputs the
real reference functions produced for them here (numpy 2.3.5 / scipy 1.18.1 / scikit-learn
1.9.0; the reference pins 2.3.3 / 1.16.2 / 1.7.2 -- same algorithm, last-ulp differences only).
Test infrastructure only.
"""

from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import scipy.sparse as sp

from oracle import real_reference as rr
from oracle.reference_paths import dict_to_arrays
from tvbingefriend_recommendation_service_b200.synthetic import Catalogue, make_catalogue

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def pack_catalogue(cat: Catalogue) -> dict:
    t = sp.csr_matrix(cat.text_features)
    return {
        "genre": cat.genre_features, "text_data": t.data, "text_indices": t.indices,
        "text_indptr": t.indptr, "text_shape": np.asarray(t.shape, dtype=np.int64),
        "platform": cat.platform_features, "type": cat.type_features,
        "language": cat.language_features, "show_ids": cat.show_ids,
    }


def unpack_catalogue(z) -> Catalogue:
    text = sp.csr_matrix((z["text_data"], z["text_indices"], z["text_indptr"]),
                         shape=tuple(int(x) for x in z["text_shape"]))
    return Catalogue(genre_features=z["genre"], text_features=text, platform_features=z["platform"],
                     type_features=z["type"], language_features=z["language"], show_ids=z["show_ids"])


def _populate_case(name: str, cat: Catalogue, cases: list[dict]) -> None:
    out = pack_catalogue(cat)
    out["n_cases"] = np.asarray(len(cases))
    for c, kw in enumerate(cases):
        captured, stats = rr.run_populate(cat, **kw)
        k = kw.get("top_n_per_show", 20)
        idx, cnt, sc = dict_to_arrays(captured, cat.show_ids.tolist(), k)
        out[f"case{c}_params"] = np.asarray([kw.get("genre_weight", 0.4), kw.get("text_weight", 0.5),
                                             kw.get("metadata_weight", 0.1), k,
                                             kw.get("min_similarity", 0.1)], dtype=np.float64)
        out[f"case{c}_idx"] = idx.astype(np.int32)
        out[f"case{c}_cnt"] = cnt.astype(np.int32)
        out[f"case{c}_scores"] = sc
        out[f"case{c}_total_records"] = np.asarray(stats["total_records"])
        print(f"{name} case{c} {kw}: {stats['total_records']} records, {len(captured)} shows")
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)


def make_c1() -> None:
    """BASELINE.json configs[0]: 1 000 shows x 5 000 vocab through the UNMODIFIED production loop,
    all rows, reference defaults (top-20, min_similarity 0.1, weights 0.4/0.5/0.1)."""
    from tvbingefriend_recommendation_service_b200.synthetic import make_config

    _populate_case("populate_c1", make_config("C1"), [{}])


def make_v500() -> None:
    """The reference's production shape in miniature: ``max_text_features`` = 500
    (scripts/compute_features.py:174), metadata widths 21 / 5 / 6, 1 500 shows through the UNMODIFIED
    production loop, all rows, reference defaults and the raw-weight case.  On the GPU this is the
    regime of the folded operand (genre / metadata as K columns of the text GEMM)."""
    cat = make_catalogue(1500, 500, nnz=20, n_genres=40, meta=(21, 5, 6), seed=2026)
    _populate_case("populate_v500_n1500", cat, [
        {},
        {"genre_weight": 21.0, "text_weight": 5.0, "metadata_weight": 6.0, "top_n_per_show": 10, "min_similarity": 2.0},
    ])


def main() -> None:
    logging.disable(logging.CRITICAL)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    import sys

    if "--only-c1" in sys.argv:
        make_c1()
        return
    if "--only-v500" in sys.argv:
        make_v500()
        return
    make_c1()
    make_v500()

    # (1) production loop (variant B) on a synthetic catalogue
    cat = make_catalogue(300, 2000, nnz=30, n_genres=40, meta=(5, 3, 2), seed=7)
    _populate_case("populate_n300", cat, [
        {},
        {"top_n_per_show": 7, "min_similarity": 0.5},
        {"genre_weight": 2.0, "text_weight": 3.0, "metadata_weight": 1.0,
         "top_n_per_show": 10, "min_similarity": 0.1},
        {"genre_weight": 0.5, "text_weight": 0.5, "metadata_weight": 0.0,
         "top_n_per_show": 50, "min_similarity": 0.0},
    ])

    # (2) the reference test's own style of input: dense random floats in every group
    #     (tests/test_scripts/test_populate_database.py:176-177 feeds np.random.rand)
    rng = np.random.default_rng(11)
    n = 48
    cat_f = Catalogue(genre_features=rng.random((n, 5)),
                      text_features=sp.csr_matrix(rng.random((n, 100))),
                      platform_features=rng.random((n, 5)), type_features=rng.random((n, 5)),
                      language_features=rng.random((n, 5)),
                      show_ids=np.arange(1, n + 1, dtype=np.int64))
    _populate_case("populate_random_float_n48", cat_f, [{}, {"top_n_per_show": 5, "min_similarity": 0.8}])

    # (3) SimilarityComputer (variant A) and the service matrix path (variant C)
    cat_a = make_catalogue(64, 500, nnz=12, n_genres=12, meta=(4, 3, 2), seed=3)
    SC = rr.similarity_computer_class()
    out = pack_catalogue(cat_a)
    for w, weights in enumerate([(0.4, 0.5, 0.1), (2.0, 3.0, 1.0)]):
        comp = SC(*weights)
        sims = comp.compute_all_similarities(cat_a.features())
        for key, mat in sims.items():
            out[f"w{w}_{key}"] = mat
            st = comp.get_similarity_statistics(mat)
            out[f"w{w}_{key}_stats"] = np.asarray([st["mean"], st["std"], st["min"], st["max"], st["median"]])
        out[f"w{w}_weights"] = np.asarray(weights)
        queries = cat_a.show_ids[:16].tolist() + [10 ** 9]  # last id unknown -> []
        for tag, (nn, ms) in {"a": (10, 0.0), "b": (5, 0.35)}.items():
            recs = rr.run_service_matrix(sims, cat_a.show_ids.tolist(), queries, n=nn, min_similarity=ms,
                                         genre_weight=weights[0], text_weight=weights[1],
                                         metadata_weight=weights[2])
            pos = {int(s): i for i, s in enumerate(cat_a.show_ids.tolist())}
            idx = np.full((len(queries), nn), -1, dtype=np.int32)
            sc = np.full((4, len(queries), nn), np.nan)
            cnt = np.zeros(len(queries), dtype=np.int32)
            for qi, q in enumerate(queries):
                cnt[qi] = len(recs[int(q)])
                for c, rrec in enumerate(recs[int(q)]):
                    idx[qi, c] = pos[int(rrec["show_id"])]
                    sc[:, qi, c] = [rrec["similarity_score"], rrec["genre_score"],
                                    rrec["text_score"], rrec["metadata_score"]]
            out[f"w{w}_svc{tag}_params"] = np.asarray([nn, ms])
            out[f"w{w}_svc{tag}_idx"], out[f"w{w}_svc{tag}_cnt"], out[f"w{w}_svc{tag}_scores"] = idx, cnt, sc
        out["svc_queries"] = np.asarray(queries, dtype=np.int64)
    np.savez_compressed(GOLDEN / "similarity_computer_n64.npz", **out)
    print("wrote", sorted(p.name for p in GOLDEN.glob("*.npz")))


if __name__ == "__main__":
    main()
