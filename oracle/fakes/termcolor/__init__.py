"""termcolor stand-in (test infrastructure only): the reference's franka_robot.py prints coloured banners."""


def cprint(*args, **kwargs):
    pass
