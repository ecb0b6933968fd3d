"""Name -> class registry and build_network, with the reference's semantics
(basicsr/utils/registry.py:4-79, basicsr/archs/__init__.py:19-25): `register()` stores `cls.__name__`,
registering a duplicate name asserts, `get` of an unknown name raises KeyError, and
`build_network(opt)` deep-copies the option dict, pops 'type' and instantiates with the rest as kwargs.
"""
import logging
from copy import deepcopy


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, (f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())

    def keys(self):
        return self._obj_map.keys()


ARCH_REGISTRY = Registry('arch')


def build_network(opt):
    opt = deepcopy(opt)
    network_type = opt.pop('type')
    net = ARCH_REGISTRY.get(network_type)(**opt)
    logging.getLogger('basicsr').info(f'Network [{net.__class__.__name__}] is created.')
    return net
