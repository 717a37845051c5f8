import os

data_dir = os.path.join(os.path.dirname(__file__), "data")
