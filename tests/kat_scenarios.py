"""The reference's own known-answer tests for this path, restated as data.

Each scenario cites the reference test it restates (paths relative to /root/reference/tests).  The same data
drives tests/test_reference_kats.py (oracle, CPU) and tests/test_gpu_reference_kats.py (CUDA path, C ABI)."""
import numpy as np

# ---- test_utils.py:19-27 test_rotated_rectangles_intersect: ((center, length, width, angle) x 2, expected)
RECTANGLES = [
    (([12.86076812, 28.60182391], 5.0, 2.0, -0.4675779906495494), ([9.67753944, 28.90585412], 5.0, 2.0, -0.3417019364473201), True),
    (([0, 0], 2, 1, 0), ([0, 1], 2, 1, 0), True),
    (([0, 0], 2, 1, 0), ([0, 2.1], 2, 1, 0), False),
    (([0, 0], 2, 1, 0), ([1, 1.1], 2, 1, 0), False),
    (([0, 0], 2, 1, np.pi / 4), ([1, 1.1], 2, 1, 0), True),
]


# ---- road/test_neighbour_vehicles.py: fixtures :39-148 as builders over the product's NetworkTable
def _straight(net, a, b, x0, x1, y=0.0):
    net.add_straight(a, b, [x0, y], [x1, y])


def straight_connected_road(net):      # :39-58  a->b (50 m), b->c (50 m)
    _straight(net, "a", "b", 0, 50)
    _straight(net, "b", "c", 50, 100)


def straight_curve_road(net):          # :61-84  a->b then a circular arc b->c
    _straight(net, "a", "b", 0, 50)
    net.add_circular("b", "c", [50, -20], 20, np.deg2rad(90), np.deg2rad(0), clockwise=False)


def three_segment_road(net):           # :87-115
    _straight(net, "a", "b", 0, 50)
    _straight(net, "b", "c", 50, 100)
    _straight(net, "c", "d", 100, 150)


def multi_lane_road(net):              # :118-148 two lanes per segment
    _straight(net, "a", "b", 0, 50, 0)
    _straight(net, "a", "b", 0, 50, 4)
    _straight(net, "b", "c", 50, 100, 0)
    _straight(net, "b", "c", 50, 100, 4)


def only_ab(net):                      # :303-311
    _straight(net, "a", "b", 0, 50)


def only_bc(net):                      # :323-331
    _straight(net, "b", "c", 50, 100)


# (name, reference lines, road builder, connected flag, vehicles [(lane_index, longitudinal)], query lane of
#  vehicle 0 = ego, expected front, expected rear) — expected: index into the vehicle list, None, or "any"
NEIGHBOURS = [
    ("front_and_rear_on_same_segment", "159-167", straight_connected_road, False,
     [(("a", "b", 0), 25), (("a", "b", 0), 40), (("a", "b", 0), 10)], ("a", "b", 0), 1, 2),
    ("no_neighbours", "169-175", straight_connected_road, False, [(("a", "b", 0), 25)], ("a", "b", 0), None, None),
    ("only_front", "177-184", straight_connected_road, False,
     [(("a", "b", 0), 10), (("a", "b", 0), 40)], ("a", "b", 0), 1, None),
    ("only_rear", "186-193", straight_connected_road, False,
     [(("a", "b", 0), 40), (("a", "b", 0), 10)], ("a", "b", 0), None, 1),
    ("connected_segments_ignored_by_default", "195-202", straight_connected_road, False,
     [(("a", "b", 0), 48), (("b", "c", 0), 5)], ("a", "b", 0), None, None),
    ("front_on_next_segment", "213-224", straight_connected_road, True,
     [(("a", "b", 0), 48), (("b", "c", 0), 5)], ("a", "b", 0), 1, "any"),
    ("rear_on_previous_segment", "226-237", straight_connected_road, True,
     [(("b", "c", 0), 5), (("a", "b", 0), 45)], ("b", "c", 0), "any", 1),
    ("front_on_curve_segment", "239-250", straight_curve_road, True,
     [(("a", "b", 0), 48), (("b", "c", 0), 5)], ("a", "b", 0), 1, "any"),
    ("closer_same_segment_preferred", "252-269", straight_connected_road, True,
     [(("a", "b", 0), 30), (("a", "b", 0), 45), (("b", "c", 0), 10)], ("a", "b", 0), 1, "any"),
    ("both_connected_front_and_rear", "271-281", three_segment_road, True,
     [(("b", "c", 0), 5), (("a", "b", 0), 45), (("c", "d", 0), 5)], ("b", "c", 0), 2, 1),
    ("multi_lane_same_lane_id", "283-297", multi_lane_road, True,
     [(("a", "b", 0), 48), (("b", "c", 0), 5), (("b", "c", 1), 3)], ("a", "b", 0), 1, "any"),
    ("no_next_segment", "303-321", only_ab, True, [(("a", "b", 0), 48)], ("a", "b", 0), None, None),
    ("no_previous_segment", "323-341", only_bc, True, [(("b", "c", 0), 5)], ("b", "c", 0), None, None),
    ("vehicle_far_on_next_segment_detected", "343-352", straight_connected_road, True,
     [(("a", "b", 0), 25), (("b", "c", 0), 40)], ("a", "b", 0), 1, "any"),
]


def build_neighbour_case(case):
    """-> (NetworkTable, x, y, heading, lane index per vehicle, query lane index)"""
    from highwayenv_b200.road.network import NetworkTable

    _, _, builder, _, vehicles, query, _, _ = case
    net = NetworkTable()
    builder(net)
    net.finalize()
    xs, ys, hs, lanes = [], [], [], []
    for lane_index, lon in vehicles:  # _make_vehicle (:14-29): on the lane centre, lane_index forced
        l = net.index[lane_index]
        px, py = net.position(l, float(lon), 0.0)
        xs.append(float(px)), ys.append(float(py)), hs.append(float(net.heading_at(l, float(lon)))), lanes.append(l)
    return net, np.array(xs), np.array(ys), np.array(hs), np.array(lanes, dtype=np.int32), net.index[query]


def check_neighbour_result(case, front, rear):
    name, _, _, _, _, _, e_front, e_rear = case
    for got, want, what in ((front, e_front, "front"), (rear, e_rear, "rear")):
        if want == "any":
            continue
        assert got == (-1 if want is None else want), f"{name}: {what} = {got}, reference expects {want}"


FPS = 15  # vehicle/test_dynamics.py:9, vehicle/test_control.py:8


def diamond_network():
    """road/test_road.py:9-20: the diamond 0->1->2->0 / 1->3->0 (node names as strings here)"""
    from highwayenv_b200.road.network import NetworkTable

    net = NetworkTable()
    net.add_straight("0", "1", [0, 0], [10, 0])
    net.add_straight("1", "2", [10, 0], [5, 5])
    net.add_straight("2", "0", [5, 5], [0, 0])
    net.add_straight("1", "3", [10, 0], [5, -5])
    net.add_straight("3", "0", [5, -5], [0, 0])
    net.finalize()
    return net


def single_lane_road():
    """RoadNetwork.straight_road_network(lanes=1) (road/road.py:291-321): one 10 km lane "0" -> "1\""""
    from highwayenv_b200.road.network import NetworkTable

    net = NetworkTable()
    net.add_straight("0", "1", [0, 0], [10000, 0], speed_limit=30.0)
    net.finalize()
    return net
