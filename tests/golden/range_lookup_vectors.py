"""Vectors of bifromq-dist/bifromq-dist-server/src/test/java/org/apache/bifromq/dist/server/scheduler/TenantRangeLookupCacheTest.java
(:109-330), transcribed: (topic, ordered candidates, kept candidate indices). A candidate is None (no Fact) or (first, last) global
filter levels (either None when the Fact lacks it); level 0 is the tenant id."""
T = "tenantA"


def F(*levels):
    return [T] + list(levels)


VECTORS = [
    ("a", [], []),                                                                          # emptyRouterReturnsEmpty :109-113
    ("m/n", [(F("a"), F("z"))], [0]),                                                       # singleCandidateFullFactCovers :115-127
    ("a/b", [(F("m"), F("z"))], []),                                                        # singleCandidateFullFactNotCover :129-140
    ("a", [(F("a"), None)], []),                                                            # ...MissingFirstOrLastIsEmptyRange :142-162
    ("a", [(None, F("z"))], []),
    ("topic", [None], [0]),                                                                 # singleCandidateNoFactIsIncluded :164-174
    ("n/1", [(F("a"), F("z")), (F("n"), F("s")), (F("t"), F("z"))], [0, 1]),                # multiCandidatesTwoCoveringAndOneNot :176-197
    ("z/1", [None, (F("x"), F("z")), (F("z"), F("zz"))], [0, 2]),                           # multiCandidatesMixWithNoFact :199-220
    ("n/1", [(F("a"), F("m")), (F("n"), F("z"))], [1]),                                     # ...LastLessThanTopicThenFollowingCovers :222-238
    ("a", [(F("b"), F("c")), None], []),                                                    # earlyStopTopicLessThanFirstOfFirstCandidate :240-256
    ("a", [None, (F("b"), F("c"))], [0]),                                                   # earlyStopAfterIncludingNoFactFirst :258-275
    ("b", [(F("b"), F("z"))], [0]),                                                         # includeWhenEqualToFirstOrLast :277-291
    ("b", [(F("a"), F("b"))], [0]),
    ("a/b/c", [(F("a", "b"), F("a", "z")), (F("b"), F("c"))], [0]),                          # multiLevelTopicAndOrder :293-311
    ("n/1", [(F("x"), F("z"))], []),                                                        # cacheFunctionalConsistency :313-330
]
