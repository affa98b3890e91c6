"""Parse the reference's output files (golden fixtures) into comparable Python values."""
import json


def parse_pair_csv(data: bytes):
    lines = data.decode().splitlines()
    assert lines[0] == "Agent1,Agent2,N_Contacts,TotalContactDuration"
    return [tuple(int(v) for v in ln.split(",")) for ln in lines[1:]]


def parse_agent_csv(data: bytes):
    lines = data.decode().splitlines()
    assert lines[0] == "agent_ID,Number_of_Contacts"
    out = []
    for i, ln in enumerate(lines[1:]):
        k, v = ln.split(",")
        assert k == "ID_%d" % (i + 1)
        out.append(int(v))
    return out


def parse_coord_csv(data: bytes):
    lines = data.decode().splitlines()
    assert lines[0] == "x,y,n_contacts"
    out = {}
    for ln in lines[1:]:
        x, y, c = ln.split(",")
        out[(float(x), float(y))] = int(c)
    return out


def parse_statistics(data: bytes):
    return json.loads(data.decode())["data"]


def parse_contacts_txt(data: bytes):
    """contacts.txt -> {step: [(x, y, count), ...]} in file order."""
    lines = data.decode().splitlines()
    out = {}
    i = 0
    while i < len(lines):
        n = int(lines[i])
        assert lines[i + 1].startswith("step :")
        step = int(lines[i + 1][6:])
        rows = []
        for ln in lines[i + 2 : i + 2 + n]:
            x, y, c = ln.split("\t")
            rows.append((float(x), float(y), int(c)))
        out[step] = rows
        i += 2 + n
    return out
