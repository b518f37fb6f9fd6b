"""Drop-in counterparts of the reference's ``predict_model`` package (predict.py:5-9 imports the five
constructors from here by these module and symbol names)."""


def _ctors():
    from .bam import Xception_DeepLabV3_Plus_bam
    from .hrnet import HRNet
    from .res34 import ResNetFamily
    from .scse import UNet
    from .v3plus import Xception_DeepLabV3_Plus
    # in the order predict.py:run_model (75-87) runs them
    return {"res34": lambda: ResNetFamily().run_model("res34"), "hrnet": HRNet,
            "v3plus": Xception_DeepLabV3_Plus, "scse": lambda: UNet(2), "bam": Xception_DeepLabV3_Plus_bam}


class _Lazy(dict):
    def __missing__(self, key):
        self.update(_ctors())
        return dict.__getitem__(self, key)

    def __iter__(self):
        if not len(self):
            self.update(_ctors())
        return dict.__iter__(self)


CTORS = _Lazy()
MODEL_NAMES = ("res34", "hrnet", "v3plus", "scse", "bam")
