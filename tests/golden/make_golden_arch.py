"""Generates tests/golden/arch_b0_blocks.json by EXECUTING the reference's own architecture code.  Run in the build
container (needs /root/reference):

    python tests/golden/make_golden_arch.py

The EfficientNet builder modules import TensorFlow and cannot be imported here, but the code that derives the backbone's
block table is plain Python.  It is executed from the reference's source text (ast), nothing is restated:

* models/efficientnet/efficientnet_builder.py:29-43 `efficientnet_params`, :45-123 `BlockDecoder` (with the truncation at
  `max_block_num`, :99-109), :125-153 `efficientnet` (the seven block strings and the global parameters);
* models/efficientnet/efficientnet_model.py:42-60 `GlobalParams` / `BlockArgs`, :106-131 `round_filters` / `round_repeats`;
* models/efficientnet/efficientnet_model.py:329-349: the `for block_args in self._blocks_args` loop of `Model._build`
  (filter rounding, repeat expansion, stride 1 and input = output filters for the repeats), run on a stand-in `self` whose
  block class records the `BlockArgs` it is constructed with.

`tf` is a stub that only provides `tf.nn.swish` (a value stored in GlobalParams) and a silent `tf.logging.info`.
"""
import ast
import collections
import json
import math
import os
import re
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/models/efficientnet"


def _load(path, names, assigns=()):
    tree = ast.parse(open(path).read())
    body = []
    for n in tree.body:
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names:
            body.append(n)
        elif isinstance(n, ast.Assign) and any(getattr(t, "id", None) in assigns for t in n.targets):
            body.append(n)
        elif (isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute)
              and getattr(n.targets[0].value, "value", None) is not None
              and getattr(n.targets[0].value.value, "id", None) in assigns):      # X.__new__.__defaults__ = ...
            body.append(n)
    return ast.Module(body=body, type_ignores=[]), tree


def main():
    tf = types.SimpleNamespace(nn=types.SimpleNamespace(swish="swish"),
                               logging=types.SimpleNamespace(info=lambda *a, **k: None))
    ns = {"collections": collections, "math": math, "re": re, "tf": tf, "np": None, "xrange": range}
    mod, model_tree = _load(os.path.join(REF, "efficientnet_model.py"), {"round_filters", "round_repeats"},
                            {"GlobalParams", "BlockArgs"})
    exec(compile(mod, "ref_efficientnet_model", "exec"), ns)
    ns["efficientnet_model"] = types.SimpleNamespace(BlockArgs=ns["BlockArgs"], GlobalParams=ns["GlobalParams"])
    mod, _ = _load(os.path.join(REF, "efficientnet_builder.py"), {"efficientnet_params", "BlockDecoder", "efficientnet"})
    exec(compile(mod, "ref_efficientnet_builder", "exec"), ns)
    # the block-expansion loop of Model._build, verbatim
    model_cls = next(n for n in model_tree.body if isinstance(n, ast.ClassDef) and n.name == "Model")
    build = next(n for n in model_cls.body if isinstance(n, ast.FunctionDef) and n.name == "_build")
    loop = next(n for n in build.body if isinstance(n, ast.For))
    loop_code = compile(ast.Module(body=[loop], type_ignores=[]), "ref_Model__build_loop", "exec")

    out = {"efficientnet_params": {}}
    w, d, res, drop = ns["efficientnet_params"]("efficientnet-b0")
    out["efficientnet_params"]["efficientnet-b0"] = [w, d, res, drop]
    for max_block_num in (10, None):
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):          # the decoder prints when it truncates
            blocks_args, gp = ns["efficientnet"](w, d, drop, max_block_num=max_block_num)

        class Recorder:
            def __init__(self, block_args, global_params):
                self.args = block_args

        self_ = types.SimpleNamespace(_blocks_args=blocks_args, _global_params=gp, _blocks=[],
                                      _get_conv_block=lambda conv_type: Recorder)
        scope = dict(ns)
        scope["self"] = self_
        exec(loop_code, scope)
        rows = []
        for b in self_._blocks:
            a = b.args
            rows.append({"kernel_size": a.kernel_size, "strides": list(a.strides), "input_filters": a.input_filters,
                         "output_filters": a.output_filters, "expand_ratio": a.expand_ratio, "se_ratio": a.se_ratio,
                         "id_skip": bool(a.id_skip), "conv_type": a.conv_type})
        out["max_block_num_%s" % max_block_num] = {
            "blocks": rows,
            "global_params": {"batch_norm_momentum": gp.batch_norm_momentum, "batch_norm_epsilon": gp.batch_norm_epsilon,
                              "drop_connect_rate": gp.drop_connect_rate, "depth_divisor": gp.depth_divisor,
                              "width_coefficient": gp.width_coefficient, "depth_coefficient": gp.depth_coefficient}}
    path = os.path.join(HERE, "arch_b0_blocks.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: len(v["blocks"]) for k, v in out.items() if k.startswith("max")})


if __name__ == "__main__":
    main()
