class RecordVideo:  # only referenced in type hints / optional hooks
    pass
