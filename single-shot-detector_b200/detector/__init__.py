from .ssd import SSD
