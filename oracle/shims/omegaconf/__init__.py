class OmegaConf:  # stand-in; htdemucs configs are out of scope
    @staticmethod
    def load(path):
        raise RuntimeError("omegaconf is not installed")
