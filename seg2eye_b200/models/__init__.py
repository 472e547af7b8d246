"""Drop-in for the reference's `models` package (models/__init__.py)."""
import importlib

import torch


def find_model_using_name(model_name):
    modellib = importlib.import_module(__name__ + "." + model_name + "_model")
    target = model_name.replace('_', '') + 'model'
    for name, cls in modellib.__dict__.items():
        if name.lower() == target.lower() and isinstance(cls, type) and issubclass(cls, torch.nn.Module):
            return cls
    raise ValueError("no model class matching %s" % target)


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % (type(instance).__name__))
    return instance
