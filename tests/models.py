"""The reference's own models (tests/testthat/*.R, README.md, vignettes/*.Rmd of /root/reference), written
once against the modelling surface shared by the product host (easylp_b200.model) and the dense assembly
oracle (oracle.dsl_ref).  `api` is either module; each builder returns the model object.

Where R writes  `for (i in I) expr`  inside `$con`, Python writes  for_(lambda i: expr, i=I);
R's `sum(...)` is api.Sum(...); `x[, m]` is x[:, m].  Index sets keep R's 1-based positions / names.
"""
import numpy as np


def readme(api):
    """G1 — README.md:16-24"""
    lp = api.easylp()
    x = lp.var("x")
    y = lp.var("y")
    lp.max(x + y)
    lp.con(x + 2 * y <= 3,
           y >= 3 * x - 2)
    return lp


def dop(api):
    """G2 — tests/testthat/test-DOP.R (objective 3 985 000 - 45 000)"""
    DOP = ["Empordà", "Garrigues", "Siurana", "Terra Alta"]
    Super = ["Girona", "Lleida", "Tarragona"]
    Moli = ["A", "B"]
    P = api.parameter
    capacitat_recolleccio = P([6000, 7000, 8000, 7000], DOP)
    coeficient_extraccio = P([.25, .3, .25, .2], DOP)
    cost_tdm = P([54, 56, 60, 49, 41, 53, 54, 52], DOP, Moli, byrow=True)
    capacitat_extraccio = P([12000, 20000], Moli)
    cost_extraccio = P([78, 82], Moli)
    cost_tms = P([47, 56, 58, 51, 52, 59], Super, Moli, byrow=True)     # [Super x Molí] on purpose, like the test
    demanda = P([1500, 3000, 2500], Super)

    lp = api.easylp()
    tdm = lp.var("tdm", DOP, Moli, lower=0)
    tms = lp.var("tms", Moli, Super, lower=0)
    lp.min(api.Sum(cost_tdm * tdm)
           + api.sum_for(lambda m: tdm[:, m] * cost_extraccio[m], m=Moli)
           + api.Sum(cost_tms * tms)
           - 45000)
    lp.alias(rec=api.rowSums(tdm), ext=api.rowSums(tms))
    rec, ext = lp.aliases["rec"], lp.aliases["ext"]
    lp.con(
        tdm_ext=api.for_(lambda m: api.sum_for(lambda d: tdm[d, m] * coeficient_extraccio[d], d=DOP) == ext[m], m=Moli),
        recolleccio=api.for_(lambda d: rec[d] <= capacitat_recolleccio[d], d=DOP),
        extraccio=api.for_(lambda m: api.Sum(tdm[:, m]) <= capacitat_extraccio[m], m=Moli),
        satisfaccio=api.for_(lambda s: api.Sum(tms[:, s]) >= demanda[s], s=Super),
    )
    return lp


def unbounded(api):
    """G3 — tests/testthat/test-unbounded.R"""
    lp = api.easylp()
    x = lp.var("x")
    lp.max(x)
    return lp


def rhs_variable(api):
    """G4 — vignettes/constraints.Rmd:225-230: `2 >= x` is stored as  -x >= -2"""
    lp = api.easylp()
    x = lp.var("x")
    lp.con(api.compare(2, ">=", x))
    return lp


def infeasible_mean(api):
    """G5 — vignettes/constraints.Rmd:313-334"""
    lp = api.easylp()
    x = lp.var("x", range(1, 5))
    lp.min(api.Sum(x))
    lp.con(limit=x <= 2, average=api.mean(x) >= 3)
    return lp


def constraints(api):
    """G6 — tests/testthat/test-constraints.R:1-20"""
    A, B, C = [1, 2], [1, 2, 3], [1, 2]
    lp = api.easylp()
    x = lp.var("x", A, B, C)
    y = lp.var("y", B)
    z = lp.var("z", A, B, C)
    lp.con(
        r1=api.for_(lambda b: api.Sum(x[:, b, :]) <= y[b], b=B),
        r2=api.for_(lambda a: api.for_(lambda b: x[a, b, 1] >= y[b] / 2 + 1, b=B), a=A),
        r3=api.for_(lambda b: x[:, b, 2] >= 1, b=B),
        r4=x <= z,
        r5=api.cumsum(2 * y + 1) >= 0,
        r6=-x > 2,
    )
    lp.uncon("r3")
    return lp


def forsplit(api):
    """G6 — tests/testthat/test-forsplit.R: triangular nested for"""
    lp = api.easylp()
    x = lp.var("x", range(1, 5), range(1, 5))
    lp.con(hi=api.for_(lambda i: api.for_(lambda j: x[i, j] == 1, j=range(i, 5)), i=range(1, 5)))
    return lp


def aliases(api):
    """G6 — tests/testthat/test-aliases.R"""
    factory, market = ["A", "B"], [1, 2]
    capacity = api.parameter([120, 180], factory)
    demand = api.parameter([140, 150], market)
    lp = api.easylp()
    t = lp.var("t", factory, market, lower=0)
    lp.alias(Fac=factory, Mar=market, made=api.rowSums(t), sold=api.colSums(t))
    made, sold = lp.aliases["made"], lp.aliases["sold"]
    lp.con(cap=api.for_(lambda i: made[i] <= capacity[i], i=lp.aliases["Fac"]),
           dem=api.for_(lambda j: sold[j] >= demand[j], j=lp.aliases["Mar"]))
    return lp


def transport_vignette(api, sum_for_objective=False):
    """G7 — vignettes/easylp.Rmd:41-60 (continuous relaxation; data integral) and :135 (sum_for objective)"""
    factory, market = ["A", "B", "C"], [1, 2, 3, 4]
    supply = api.parameter([50, 30, 45], factory)
    demand = api.parameter([30, 25, 40, 15], market)
    cost = api.parameter([51, 89, 64, 32, 28, 87, 66, 48, 82, 78, 66, 29], factory, market, byrow=True)
    lp = api.easylp()
    x = lp.var("x", factory, market, lower=0)
    if sum_for_objective:
        lp.min(api.sum_for(lambda f, m: cost[f, m] * x[f, m], f=factory, m=market))
    else:
        lp.min(api.Sum(cost * x))
    lp.con(make=api.for_(lambda f: api.Sum(x[f, :]) <= supply[f], f=factory),
           sell=api.for_(lambda m: api.Sum(x[:, m]) >= demand[m], m=market))
    return lp


def modified(api, seed=0):
    """tests/testthat/test-modified.R:2-14 (the reference draws an unseeded runif objective; seeded here)"""
    rng = np.random.default_rng(seed)
    lp = api.easylp()
    x = lp.var("x", [1, 2, 3], [1, 2, 3], lower=1, upper=10)
    y = lp.var("y", [1, 2], [1, 2], [1, 2], lower=1, upper=10)
    lp.min(api.Sum(x * rng.uniform(-1, 1, 9)) + api.Sum(y * rng.uniform(-1, 1, 8)))
    lp.con(api.rowSums(x) == api.colSums(x),
           api.diag(x)[[2, 3]] == [1, 2],
           api.apply(y, [1, 2], api.mean) == [2, 3, 4, 5])
    return lp


def modified_indexed(api):
    """tests/testthat/test-modified.R:24-41"""
    lp = api.easylp()
    x = lp.var("x", d1=list("abcd"), d2=list("ABC"), d3=[1, 2], lower=-10, upper=10)
    lp.min(api.Sum(x))
    lp.con(api.rowSums(x)[1] == 3,
           api.rowSums(x)["b"] == 4,
           api.apply(x, [1, 2], api.mean)[[1, 2], "B"] == 2)
    return lp


def cyingair(api):
    """tests/testthat/test-cyingair.R:1-24 — binary + integer variables, `$associate`; the reference's test pins the
    solution x = (0, 2, 3, 49), quin = (0, 1, 1, 1) (test-cyingair.R:27-30)"""
    Avio = ["Jumbo", "Petit", "Mitja", "Gran"]
    preu = [79, 67, 50, 35]
    benefici = [5.8, 4.2, 3, 2.3]
    lp = api.easylp()
    quin = lp.var("quin", Avio, binary=True)
    x = lp.var("x", Avio, integer=True, lower=0, upper=100)
    lp.max(api.Sum(x * benefici))
    lp.associate(x, quin, min1=1)
    lp.con(tipus=api.Sum(quin) == 3,
           r_pressupost=api.Sum(x * preu) <= 2000,
           min_avions=api.Sum(x) >= 35,
           no_mes_petits_que_mitjans=x["Petit"] <= x["Mitja"],
           no_jumbo_i_grans=quin["Jumbo"] + quin["Gran"] <= 1,
           quinze_percent=x["Jumbo"] <= 0.15 * api.Sum(x))
    return lp


def investments_assembly(api):
    """tests/testthat/test-investments.R:1-41 — binary variables; the reference's test pins objective 469 and
    x = (0, 0, 1, 1, 1, 0) (test-investments.R:45-46)"""
    Project, Year = list(range(1, 7)), list(range(1, 6))
    npv = api.parameter([141, 187, 121, 83, 265, 127], Project)
    budget = api.parameter([250, 75, 50, 50, 50], Year)
    investment = api.parameter([75, 25, 20, 15, 10, 90, 35, 0, 0, 30, 60, 15, 15, 15, 15,
                                30, 20, 10, 5, 5, 100, 25, 20, 20, 20, 50, 20, 10, 30, 40], Project, Year, byrow=True)
    na = np.nan
    incompatible = api.parameter([na, 1, 0, 1, 0, 0, na, na, 1, 0, 0, 0, na, na, na, 0, 0, 0,
                                  na, na, na, na, 0, 0, na, na, na, na, na, 1, na, na, na, na, na, na],
                                 Project, Project, byrow=True)
    lp = api.easylp()
    x = lp.var("x", Project, binary=True)
    lp.max(api.Sum(x * npv))
    lp.con(
        budget=api.for_(lambda a: api.sum_for(lambda p: x[p] * investment[p, a], p=Project) <= budget[a], a=Year),
        compatibility=api.for_(lambda p: api.for_(lambda q: x[p] + x[q] + incompatible[p, q] <= 2,
                                                  q=range(p + 1, len(Project) + 1)), p=Project[:-1]),
    )
    return lp


def duplicate_fold(api, seed=3):
    """Not in the reference's tests: a sum_for whose grid hits the same column many times with values whose
    sum depends on the order — exercises the ordered left fold (R/methods.R:248-250) and colSums."""
    rng = np.random.default_rng(seed)
    n = 7
    w = rng.normal(size=(n, 40)) * 10.0 ** rng.integers(-9, 9, size=(n, 40))
    lp = api.easylp()
    x = lp.var("x", range(1, n + 1), lower=0)
    lp.min(api.Sum(x))
    lp.con(
        fold=api.for_(lambda i: api.sum_for(lambda k: x[1 + (i + k) % n] * w[i - 1, k - 1] + x[i] * w[(i + k) % n, k - 1],
                                            k=range(1, 41)) >= 1.0, i=range(1, n + 1)),
        colsum=api.Sum(api.cumsum(x * w[:, 0]) + x / 3) <= 7,
        mixed=api.Sum(x[[1, 2, 3]] * w[0, :3], x[[2, 3, 4]] * w[1, :3], 5.0, x[[1, 2, 3]] / 7) == 1,
    )
    return lp


ALL = dict(readme=readme, dop=dop, unbounded=unbounded, rhs_variable=rhs_variable, infeasible_mean=infeasible_mean,
           constraints=constraints, forsplit=forsplit, aliases=aliases, transport_vignette=transport_vignette,
           transport_sum_for=lambda api: transport_vignette(api, True), modified=modified,
           modified_indexed=modified_indexed, investments_assembly=investments_assembly, duplicate_fold=duplicate_fold,
           cyingair=cyingair)
