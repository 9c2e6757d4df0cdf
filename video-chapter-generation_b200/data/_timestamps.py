"""Host helpers of the datasets: the chapter CSV and the "MM:SS title" timestamp strings.
Same names / results as the reference's data/common_utils.py (parse_csv_to_list :6-15, extract_timestamp :37-68,
extract_first_timestamp :71-83); written from the behaviour, not the text."""
import re

TIMESTAMP_DELIMITER = "%^&*"
# the reference tries these in this order and takes the first pattern that occurs anywhere in the string
_PATTERNS = [re.compile(p) for p in (r"\d{2}:\d{2}:\d{2}", r"\d{1}:\d{2}:\d{2}", r"\d{2}:\d{2}", r"\d{1}:\d{2}")]


def parse_csv_to_list(csv_file):
    """-> (videoId list, title list, duration list, list of per-video timestamp-string lists)."""
    import pandas as pd
    table = pd.read_csv(csv_file)
    stamps = [str(x).split(TIMESTAMP_DELIMITER) for x in table["timestamp"].values]
    return list(table["videoId"].values), list(table["title"].values), list(table["duration"].values), stamps


def extract_timestamp(s):
    """-> (matched text, seconds, start index, end index), or ("", -1, -1, -1) without a timestamp."""
    for pat in _PATTERNS:
        m = pat.search(s)
        if m:
            si, ei = m.span()
            sec = 0
            for unit, field in zip((1, 60, 3600), reversed(s[si:ei].split(":"))):
                sec += unit * int(field)
            return s[si:ei], sec, si, ei
    return "", -1, -1, -1


def extract_first_timestamp(s):
    """Smallest timestamp of the string (seconds) and the string with every timestamp removed."""
    _, sec, si, ei = extract_timestamp(s)
    earliest, rest = sec, s[:si] + s[ei:]
    while sec != -1:
        _, sec, si, ei = extract_timestamp(rest)
        if sec != -1:
            earliest = min(earliest, sec)
            rest = rest[:si] + rest[ei:]
    return earliest, rest
