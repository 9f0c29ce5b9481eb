from .act import Activation1d
from .resample import UpSample1d, DownSample1d
from .filter import LowPassFilter1d, kaiser_sinc_filter1d
