def register_keras_serializable(package="Custom", name=None):
    def deco(cls):
        return cls
    return deco
