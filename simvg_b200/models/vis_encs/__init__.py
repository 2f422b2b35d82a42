from .beit.beit3 import BEIT3
