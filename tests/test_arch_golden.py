"""The backbone's block table against the reference's OWN architecture code: tests/golden/arch_b0_blocks.json is written
by tests/golden/make_golden_arch.py, which executes efficientnet_builder.py (block strings, BlockDecoder with the
max_block_num truncation), efficientnet_model.py round_filters / round_repeats and the block-expansion loop of
Model._build from the reference's source text.  The oracle's table must equal it; tests/test_abi.py ties the C library's
parameter table to the oracle's, so the engine's layer shapes are pinned to the reference transitively."""
import json
import os

from oracle import efficientlab_oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "arch_b0_blocks.json")))


def test_efficientnet_b0_coefficients():
    assert GOLD["efficientnet_params"]["efficientnet-b0"] == [1.0, 1.0, 224, 0.2]


def test_oracle_block_table_equals_the_reference_builder_output():
    ref = GOLD["max_block_num_10"]          # EfficientLab-6-3: efficientnet-b0 truncated at block 10 (efficientlab.py:134-139)
    blocks = O.decode_blocks(10)
    assert len(blocks) == len(ref["blocks"]) == 11
    gp = ref["global_params"]
    for i, (b, r) in enumerate(zip(blocks, ref["blocks"])):
        assert r["conv_type"] == 0 and r["strides"][0] == r["strides"][1]
        assert (b.kernel, b.stride, b.cin, b.cout, b.expand) == (
            r["kernel_size"], r["strides"][0], r["input_filters"], r["output_filters"], r["expand_ratio"]), i
        # efficientnet_model.py:203-204: max(1, int(input_filters * se_ratio))
        assert b.se_reduced == max(1, int(r["input_filters"] * r["se_ratio"])), i
        # efficientnet_model.py:281-287: identity skip iff id_skip, all strides 1 and input == output filters
        assert b.skip == (r["id_skip"] and r["strides"] == [1, 1] and r["input_filters"] == r["output_filters"]), i
        # efficientnet_model.py:426-428: drop_connect_rate * idx / len(blocks)
        assert abs(b.dc_rate - gp["drop_connect_rate"] * float(i) / len(blocks)) < 1e-12, i
    assert (O.BN_MOMENTUM, O.BN_EPS, O.DROP_CONNECT_RATE) == (gp["batch_norm_momentum"], gp["batch_norm_epsilon"],
                                                             gp["drop_connect_rate"])


def test_full_b0_has_sixteen_blocks_and_the_truncation_is_a_prefix():
    full, cut = GOLD["max_block_num_None"]["blocks"], GOLD["max_block_num_10"]["blocks"]
    assert len(full) == 16 and full[:len(cut)] == cut
    assert [(b.kernel, b.stride, b.cin, b.cout) for b in O.decode_blocks(15)] == [
        (r["kernel_size"], r["strides"][0], r["input_filters"], r["output_filters"]) for r in full]
