"""Speaker names and track layouts of the reference (``core/constants.py:5-118``): pure data, needed so that
ingest and WAV packing produce byte-identical channel layouts."""

SPEAKER_NAMES = ['FL', 'FR', 'FC', 'BL', 'BR', 'SL', 'SR', 'WL', 'WR', 'TFL', 'TFR', 'TSL', 'TSR', 'TBL', 'TBR']

LEFT_SIDE_SPEAKERS = frozenset({'FL', 'SL', 'BL', 'WL', 'TFL', 'TSL', 'TBL'})
RIGHT_SIDE_SPEAKERS = frozenset({'FR', 'SR', 'BR', 'WR', 'TFR', 'TSR', 'TBR'})
CENTER_SPEAKERS = frozenset({'FC', 'LFE'})


def speaker_side(name: str) -> str:
    n = name.upper()
    if n in LEFT_SIDE_SPEAKERS:
        return 'left'
    if n in RIGHT_SIDE_SPEAKERS:
        return 'right'
    return 'center'


TRUEHD_11CH_ORDER = ['FL', 'FR', 'FC', 'BL', 'BR', 'SL', 'SR', 'TFL', 'TFR', 'TBL', 'TBR']
TRUEHD_13CH_ORDER = ['FL', 'FR', 'FC', 'BL', 'BR', 'SL', 'SR', 'TFL', 'TFR', 'TSL', 'TSR', 'TBL', 'TBR']
CHANNEL_LAYOUT_MAP = {11: TRUEHD_11CH_ORDER, 13: TRUEHD_13CH_ORDER}

SPEAKER_DELAYS = {s: 0 for s in SPEAKER_NAMES}
# SPL change relative to the middle of the head (reference core/constants.py:80-83: all zero)
IR_ROOM_SPL = {s: {'left': 0.0, 'right': 0.0} for s in SPEAKER_NAMES}

IPSILATERAL_PAIRS = (("FL", "FR"), ("SL", "SR"), ("BL", "BR"), ("TFL", "TFR"), ("TSL", "TSR"), ("TBL", "TBR"),
                     ("FC", "FC"), ("WL", "WR"))

ONSET_GROUPS = [("FL", "FR"), ("SL", "SR"), ("BL", "BR"), ("WL", "WR"), ("TFL", "TFR"), ("TSL", "TSR"),
                ("TBL", "TBR"), ("FC",)]

HESUVI_TRACK_ORDER = ['FL-left', 'FL-right', 'SL-left', 'SL-right', 'BL-left', 'BL-right', 'FC-left', 'FR-right',
                      'FR-left', 'SR-right', 'SR-left', 'BR-right', 'BR-left', 'FC-right', 'WL-left', 'WL-right',
                      'WR-left', 'WR-right', 'TFL-left', 'TFL-right', 'TFR-left', 'TFR-right', 'TSL-left',
                      'TSL-right', 'TSR-left', 'TSR-right', 'TBL-left', 'TBL-right', 'TBR-left', 'TBR-right']

HEXADECAGONAL_TRACK_ORDER = ['FL-left', 'FL-right', 'FR-left', 'FR-right', 'FC-left', 'FC-right', 'LFE-left',
                             'LFE-right', 'BL-left', 'BL-right', 'BR-left', 'BR-right', 'SL-left', 'SL-right',
                             'SR-left', 'SR-right', 'WL-left', 'WL-right', 'WR-left', 'WR-right', 'TFL-left',
                             'TFL-right', 'TFR-left', 'TFR-right', 'TSL-left', 'TSL-right', 'TSR-left',
                             'TSR-right', 'TBL-left', 'TBL-right', 'TBR-left', 'TBR-right']


def track_name(speaker: str, side: str) -> str:
    return f"{speaker}-{side}"
