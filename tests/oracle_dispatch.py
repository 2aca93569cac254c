"""Maps a parity case (oracle/cases.py) to the oracle restatement call."""
import torch

from oracle import mgdt_oracle as O


def oracle_module(name, sd, xs):
    """Returns a tensor, or (y, raw_list) for heads."""
    x = xs[0]
    if name.startswith("conv_k3s2"):
        return O.conv_bn_act(x, sd, "", 3, 2)
    if name.startswith("conv_k1"):
        return O.conv_bn_act(x, sd, "", 1, 1)
    if name.startswith("conv_k3s1"):
        return O.conv_bn_act(x, sd, "", 3, 1)
    if name == "bottleneck":
        return O.bottleneck(x, sd, "", True)
    if name == "c2f_n2":
        return O.c2f(x, sd, "", 2, True)
    if name == "c2f_n1_noshort":
        return O.c2f(x, sd, "", 1, False)
    if name == "mspa_c2f_n1":
        return O.mspa_c2f(x, sd, "", 1, True)
    if name == "mspa_c2f_n2":
        return O.mspa_c2f(x, sd, "", 2, True)
    if name == "sppf":
        return O.sppf(x, sd, "", 5)
    if name == "simfusion_4in":
        return O.simfusion_4in(xs)
    if name == "simfusion_3in":
        return O.simfusion_3in(xs, sd, "", [16, 32, 32], 32)
    if name == "simfusion_3in_allconv":
        return O.simfusion_3in(xs, sd, "", [8, 16, 24], 32)
    if name == "convnextv2_block":
        return O.convnextv2_block(x, sd, "")
    if name == "ifm":
        return O.ifm(x, sd, "", 3)
    if name == "injection":
        return O.injection(xs, sd, "", [64, 32], 1)
    if name == "injection_flag0":
        return O.injection(xs, sd, "", [64, 32], 0)
    if name == "conv_gn":
        return O.conv_gn_act(x, sd, "", 3)
    if name == "task_decomp":
        return O.task_decomposition(x, torch.nn.functional.adaptive_avg_pool2d(x, 1), sd, "")
    if name == "detect":
        raw = O.detect_head(xs, sd, "", 5, 4)
        return O.decode(raw, [8.0, 16.0], 4, 5), raw
    if name == "toodhead":
        raw = O.tood_head(xs, sd, "", 3)
        return O.decode(raw, [8.0], 16, 3), raw
    if name == "toodhead_hid128":
        raw = O.tood_head(xs, sd, "", 80)
        return O.decode(raw, [8.0], 16, 80), raw
    raise KeyError(name)
