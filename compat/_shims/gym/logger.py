MIN_LEVEL = 30


def set_level(level):
    global MIN_LEVEL
    MIN_LEVEL = level


def warn(msg, *args):
    if MIN_LEVEL <= 30:
        print("WARN: " + (msg % args if args else msg))
