"""Empty `h5py` stand-in: the reference imports it in network.py:1 but only uses it inside
Network.fromModel / save_weights, which the golden generator never calls."""
