"""Import-path shim for the reference's ``env.AttrDict`` (env.py:5-8): a dict whose keys are attributes."""


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self
