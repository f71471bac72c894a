"""CPU, build container only: `accelerate(safepy.safe.SAFE)` -- the drop-in the INTEGRATION.md describes -- against the
UNMODIFIED reference class imported from /root/reference (absent on the GPU box: skipped there).  Checks that the
graft replaces exactly the hot-path methods, that every attribute those methods read exists on a reference instance
under the same name, and that the grafted methods reach the CUDA library (and fail loudly without a device)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from ref_import import REFERENCE_ROOT, import_reference  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE_ROOT, "safepy")),
                                reason="reference checkout not available")


@pytest.fixture(scope="module")
def grafted():
    from safepy_b200 import SafeB200Mixin, accelerate
    ref = import_reference()
    return ref, accelerate(ref.SAFE), SafeB200Mixin


def test_graft_replaces_only_the_hot_path(grafted):
    ref, SAFE, mixin = grafted
    assert SAFE.__mro__[1] is mixin and SAFE.__mro__[2] is ref.SAFE
    for name in ("define_neighborhoods", "compute_pvalues", "compute_pvalues_by_randomization",
                 "compute_pvalues_by_hypergeom", "define_top_attributes", "define_domains"):
        assert getattr(SAFE, name) is getattr(mixin, name), name
        assert hasattr(ref.SAFE, name), name                    # same method names upstream
    for name in ("load_network", "load_attributes", "validate_config", "read_config", "trim_domains", "save",
                 "plot_network", "print_output_files"):
        assert getattr(SAFE, name) is getattr(ref.SAFE, name), name


def test_reference_instance_has_what_the_grafted_methods_read(grafted):
    ref, SAFE, mixin = grafted
    sf = SAFE(verbose=False)
    for attr in ("graph", "node2attribute", "attributes", "neighborhoods", "node_distance_metric",
                 "neighborhood_radius", "neighborhood_radius_type", "background", "enrichment_type",
                 "neighborhood_score_type", "multiple_testing", "num_permutations", "random_seed", "attribute_sign",
                 "enrichment_threshold", "attribute_enrichment_min_size", "attribute_unimodality_metric",
                 "attribute_distance_metric", "attribute_distance_threshold", "verbose", "ns", "pvalues_pos",
                 "pvalues_neg", "nes", "nes_binary"):
        assert hasattr(sf, attr), attr
    # defaults the standalone class mirrors (safepy_b200.safe.DEFAULTS) are the reference's
    from safepy_b200.safe import DEFAULTS
    for key in ("background", "node_distance_metric", "neighborhood_radius", "attribute_sign", "num_permutations",
                "multiple_testing", "neighborhood_score_type", "enrichment_type", "enrichment_threshold",
                "attribute_enrichment_min_size", "attribute_unimodality_metric", "attribute_distance_metric",
                "attribute_distance_threshold"):
        assert getattr(sf, key) == DEFAULTS[key], key


def test_grafted_methods_validate_like_upstream_and_reach_the_library(grafted):
    from conftest import load_golden, net_from_golden
    from safepy_b200 import synthetic as syn
    from safepy_b200._lib import SafeB200Error
    ref, SAFE, mixin = grafted
    g = load_golden("stage1_small.npz")
    sf = SAFE(verbose=False)
    sf.graph = syn.to_networkx(net_from_golden(g))
    with pytest.raises(ValueError, match="not a valid setting"):       # reference validate_config, before any GPU work
        sf.define_neighborhoods(node_distance_metric="manhattan")
    assert sf.node_distance_metric == "shortpath_weighted_layout"      # restored by the reference's validator
    from safepy_b200 import _lib
    try:
        _lib.Context().close()
        have_gpu = True
    except SafeB200Error:
        have_gpu = False
    if have_gpu:
        sf.define_neighborhoods(neighborhood_radius=float(g["r_layout"]))
        assert np.array_equal(sf.neighborhoods.words, g["nb_layout"])
    else:
        with pytest.raises(SafeB200Error, match="no CPU fallback"):    # the graft does not fall back to upstream's code
            sf.define_neighborhoods(neighborhood_radius=float(g["r_layout"]))
