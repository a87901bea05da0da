def get_event_storage():
    raise RuntimeError("training-only facility; not on the inference path")
