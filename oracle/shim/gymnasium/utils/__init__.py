class RecordConstructorArgs:
    def __init__(self, **kwargs):
        pass
