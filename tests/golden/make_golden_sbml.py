"""Writes tests/golden/sbml_cell_cycle.xml (a hand-designed 13-species cell-cycle network using every rate-law helper the
reference's generator can emit) and runs the REFERENCE's own SBML code generator on it (oracle/_ref/sbmlgen = the
reference's src/sbml/*.cpp + vendored libsbml, built by oracle/ref/sbmlgen/Makefile) to produce
tests/golden/sbml_cell_cycle_generated.txt -- the `derivative_code` of the fixture cellpop_sbml_cell_cycle.

Run where /root/reference is mounted:
    make -C oracle/ref/sbmlgen -j16 && python tests/golden/make_golden_sbml.py && python tests/golden/make_golden_cellpop.py cellpop_sbml_cell_cycle
"""
import ast
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SBMLGEN = os.path.join(ROOT, "oracle", "_ref", "sbmlgen")

# sampled variables of prior.xml, in order (the generator maps names in kinetic laws to `parameters[i]`)
VARIABLES = ["k_syn", "k_deg", "k_act", "k_inh", "variability_scale", "stdev"]
NON_SAMPLED = ["basal"]  # <non_sampled_parameter> of the experiment -> non_sampled_parameters[0]

# id: initial amount. Species without any reaction become `constant_species` (SBMLModel.cpp:109-121); TotCycE is the
# target of an assignment rule (handled outside the generated derivative).
SPECIES = {
    "CycD": 0.1, "CycE": 0.05, "CycA": 0.02, "CycB": 0.01, "E2F": 0.1, "Rb": 1.0, "pRb": 0.0, "p27": 0.5, "Cdh1": 0.9, "Cdc20": 0.02,
    "Emi1": 0.1, "CycD2": 0.0, "CycEp27": 0.0,
    "Mitogen": 1.0, "Drug": 0.25, "TotCycE": 0.0,
}
# global SBML parameters with values: names the generator cannot map to a variable or species become literals (SBMLRatelaws.cpp:203-215)
PARAMETERS = {"Kd_e2f": 0.3, "n_rb": 2.5, "km_apc": 0.05}

# (id, reactants, products, kinetic law); stoichiometry as "2 CycD"
REACTIONS = [
    ("r01", [], ["CycD"], "k_syn * Mitogen * synthcap(CycD) + basal"),
    ("r02", ["CycD"], [], "k_deg * CycD"),
    ("r03", ["Rb"], ["pRb"], "mm(k_act * 2.0, 0.2, CycD + 0.5 * CycE, Rb)"),
    ("r04", ["pRb"], ["Rb"], "k_inh * 0.5 * pRb"),
    ("r05", [], ["E2F"], "k_syn * 0.8 * hill(pRb, 0.4, n_rb)"),
    ("r06", ["E2F"], [], "k_deg * E2F * (1 + 3.0 * Rb)"),
    ("r07", [], ["CycE"], "k_syn * hill(E2F, Kd_e2f, 4)"),
    ("r08", ["CycE"], [], "k_deg * 1.5 * CycE * (1 + 2.0 * hill(CycA, 0.5, 10))"),
    ("r09", ["CycE", "p27"], ["CycEp27"], "k_act * 5.0 * CycE * p27"),
    ("r10", ["CycEp27"], ["CycE", "p27"], "k_inh * 0.3 * CycEp27"),
    ("r11", [], ["p27"], "k_syn * 0.4"),
    ("r12", ["p27"], [], "k_deg * p27 * (0.2 + 4.0 * hill(CycE, 0.6, 16))"),
    ("r13", [], ["CycA"], "k_syn * 0.7 * hill(E2F, 0.5, 2)"),
    ("r14", ["CycA"], [], "tQSSA(k_deg * 3.0, 0.1, Cdc20, CycA)"),
    ("r15", [], ["CycB"], "k_syn * 0.3 * (0.1 + pow(CycA, 1.5))"),
    ("r16", ["CycB"], [], "k_deg * CycB * (0.1 + 5.0 * Cdh1 + 3.0 * Cdc20)"),
    ("r17", [], ["Cdh1"], "mm(k_act, km_apc, 1.0, 1.0 - Cdh1)"),
    ("r18", ["Cdh1"], [], "mm(k_inh * 4.0, km_apc, CycA + CycB + 0.2 * CycE, Cdh1)"),
    ("r19", [], ["Cdc20"], "k_syn * 0.5 * hill(CycB, 0.3, 100)"),
    ("r20", ["Cdc20"], [], "k_deg * 2.0 * Cdc20 / (1 + Emi1)"),
    ("r21", [], ["Emi1"], "k_syn * 0.2 * E2F"),
    ("r22", ["Emi1"], [], "k_deg * Emi1 * (1 + CycB * 2.0)"),
    ("r23", ["2 CycD"], ["CycD2"], "k_act * 0.1 * CycD * CycD"),
    ("r24", ["CycD2"], ["2 CycD"], "k_inh * 0.2 * CycD2 * exp(-Drug)"),
]
ASSIGNMENT_RULES = [("TotCycE", "CycE + CycEp27")]

BUILTIN = {"pow": "power", "exp": "exp", "ln": "ln"}


def mathml(node) -> str:
    if isinstance(node, ast.Expression):
        return mathml(node.body)
    if isinstance(node, ast.BinOp):
        op = {ast.Add: "plus", ast.Sub: "minus", ast.Mult: "times", ast.Div: "divide"}[type(node.op)]
        return f"<apply><{op}/>{mathml(node.left)}{mathml(node.right)}</apply>"
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
        return f"<apply><minus/>{mathml(node.operand)}</apply>"
    if isinstance(node, ast.Call):
        args = "".join(mathml(a) for a in node.args)
        if node.func.id in BUILTIN:
            return f"<apply><{BUILTIN[node.func.id]}/>{args}</apply>"
        return f"<apply><ci> {node.func.id} </ci>{args}</apply>"  # a user function: AST_FUNCTION with that name
    if isinstance(node, ast.Name):
        return f"<ci> {node.id} </ci>"
    if isinstance(node, ast.Constant):
        if isinstance(node.value, int):
            return f'<cn type="integer"> {node.value} </cn>'
        return f"<cn> {node.value!r} </cn>"
    raise ValueError(ast.dump(node))


def math(expr: str) -> str:
    return '<math xmlns="http://www.w3.org/1998/Math/MathML">' + mathml(ast.parse(expr, mode="eval")) + "</math>"


def species_ref(s: str) -> str:
    parts = s.split()
    st, sid = (parts[0], parts[1]) if len(parts) == 2 else ("1", parts[0])
    return f'<speciesReference species="{sid}" stoichiometry="{st}"/>'


def function_definition(name: str, nargs: int) -> str:
    bv = "".join(f"<bvar><ci> a{i} </ci></bvar>" for i in range(nargs))
    return (f'<functionDefinition id="{name}"><math xmlns="http://www.w3.org/1998/Math/MathML"><lambda>{bv}<ci> a0 </ci></lambda></math>'
            "</functionDefinition>")


def sbml_text() -> str:
    o = ['<?xml version="1.0" encoding="UTF-8"?>', '<sbml xmlns="http://www.sbml.org/sbml/level2/version4" level="2" version="4">',
         '<model id="cell_cycle">', "<listOfFunctionDefinitions>"]
    # placeholders: the reference recognises these four names itself (SBMLRatelaws.cpp:281-345) and never evaluates the lambdas
    o += [function_definition("hill", 3), function_definition("mm", 4), function_definition("tQSSA", 4), function_definition("synthcap", 1)]
    o += ["</listOfFunctionDefinitions>", '<listOfCompartments><compartment id="cell" size="1"/></listOfCompartments>', "<listOfSpecies>"]
    for sid, amount in SPECIES.items():
        o.append(f'<species id="{sid}" name="{sid}" compartment="cell" initialAmount="{amount!r}"/>')
    o.append("</listOfSpecies><listOfParameters>")
    for pid in VARIABLES[:4] + NON_SAMPLED:
        o.append(f'<parameter id="{pid}" value="1"/>')
    for pid, v in PARAMETERS.items():
        o.append(f'<parameter id="{pid}" value="{v!r}"/>')
    o.append("</listOfParameters><listOfRules>")
    for target, expr in ASSIGNMENT_RULES:
        o.append(f'<assignmentRule variable="{target}">{math(expr)}</assignmentRule>')
    o.append("</listOfRules><listOfReactions>")
    for rid, reactants, products, law in REACTIONS:
        o.append(f'<reaction id="{rid}" reversible="false">')
        if reactants:
            o.append("<listOfReactants>" + "".join(species_ref(s) for s in reactants) + "</listOfReactants>")
        if products:
            o.append("<listOfProducts>" + "".join(species_ref(s) for s in products) + "</listOfProducts>")
        o.append(f"<kineticLaw>{math(law)}</kineticLaw></reaction>")
    o.append("</listOfReactions></model></sbml>")
    return "\n".join(o) + "\n"


def main():
    xml = os.path.join(HERE, "sbml_cell_cycle.xml")
    with open(xml, "w") as f:
        f.write(sbml_text())
    if not os.path.exists(SBMLGEN):
        sys.exit(f"{SBMLGEN} is missing: make -C oracle/ref/sbmlgen -j16 (needs /root/reference)")
    r = subprocess.run([SBMLGEN, xml, ",".join(VARIABLES), ",".join(NON_SAMPLED)], capture_output=True, text=True)
    sys.stderr.write(r.stderr)
    if r.returncode != 0:
        sys.exit("the reference's generator failed")
    text = r.stdout.replace(xml, "tests/golden/sbml_cell_cycle.xml")
    with open(os.path.join(HERE, "sbml_cell_cycle_generated.txt"), "w") as f:
        f.write(text)
    print(text)


if __name__ == "__main__":
    main()
