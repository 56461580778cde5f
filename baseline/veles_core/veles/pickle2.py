import pickle  # noqa: F401

best_protocol = pickle.HIGHEST_PROTOCOL
