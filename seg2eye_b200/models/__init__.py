"""Drop-in for the reference's `models` package: `get_option_setter(name)` / `create_model(opt)`, which
options/base_options.py:78-81 and user scripts call with --model pix2pix."""
import importlib

_MODEL_CLASSES = {"pix2pix": ("pix2pix_model", "Pix2PixModel")}


def find_model_using_name(model_name):
    try:
        module, cls = _MODEL_CLASSES[model_name.replace("_", "").lower()]
    except KeyError:
        raise ValueError("unknown --model %r (the B200 path provides: %s)" % (model_name, ", ".join(_MODEL_CLASSES)))
    return getattr(importlib.import_module("%s.%s" % (__name__, module)), cls)


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    model = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % type(model).__name__)
    return model
