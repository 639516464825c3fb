"""Small host-side containers mirroring the ones the reference's callers use."""


class AttrDict(dict):
    """dict with attribute access (same contract as blox/core.py:4-19 `AttrDict`)."""
    __setattr__ = dict.__setitem__

    def __getattr__(self, attr):
        try:
            return dict.__getitem__(self, attr)
        except KeyError:
            raise AttributeError("Attribute %r not found" % attr)

    def __getstate__(self):
        return self

    def __setstate__(self, d):
        self.update(d)


class ParamDict(AttrDict):
    """blox/utils.py:137-142: `overwrite` merges a dict of overrides in place and returns self."""

    def overwrite(self, new_params):
        for param in new_params:
            self[param] = new_params[param]
        return self
