"""state_dict templates (keys + shapes) of module cases, built from the B200 modules on CPU."""
import mgdt_yolo_b200.modules as M

_NS = {k: getattr(M, k) for k in M.__all__}


def build(ctor):
    return eval(ctor, dict(_NS))


def state_template(ctor, expect_keys=None):
    sd = build(ctor).state_dict()
    if expect_keys is not None:
        assert list(sd.keys()) == list(expect_keys), "state_dict keys differ from the reference's"
    return sd
