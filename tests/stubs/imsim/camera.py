_CAMERAS = {}


def get_camera(name):
    return _CAMERAS[name]
