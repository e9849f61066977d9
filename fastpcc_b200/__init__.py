"""fastpcc_b200 -- B200-native (sm_100a) sparse-convolution codec hot path behind FastPCC's layer API."""
__version__ = '0.1.0'
