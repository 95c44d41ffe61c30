"""A small interpreter for the JavaScript subset used by automerge-classic's mocha test files.

Purpose: extract golden fixtures (binary changes, expected patches, expected column bytes) from the
reference's own tests in this container, where no JS engine exists.  It is a fixture-generation tool
only: tests/jsfixtures/extract.py drives it, the JSON it writes lives under tests/golden/.

Supported: const/let/var (with array/object destructuring), functions and arrow functions, object /
array literals (shorthand, computed keys, spread), template strings, regex literals (opaque), calls,
`new`, member access, the usual operators, if/else, for(;;), for-of, while, break/continue/return,
throw, try/catch.  No prototypes, no classes, no getters, no generators, no async.
"""
import math
import re


class JSUndefined:
    def __repr__(self):
        return 'undefined'

    def __bool__(self):
        return False


undefined = JSUndefined()


class JSThrow(Exception):
    def __init__(self, value):
        super().__init__(str(value))
        self.value = value


class JSError(Exception):
    """A JS Error object raised by host code: kind is 'Error' | 'RangeError' | 'TypeError'."""

    def __init__(self, kind, message):
        super().__init__('%s: %s' % (kind, message))
        self.kind, self.message = kind, message


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, v):
        self.v = v


class JSRegex:
    def __init__(self, pattern, flags):
        self.pattern, self.flags = pattern, flags

    def test(self, s):
        return re.search(self.pattern, s) is not None


# ------------------------------------------------------------------ tokenizer
TOKEN_RE = re.compile(r'''
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>'(?:\\.|[^'\\])*'|"(?:\\.|[^"\\])*")
  | (?P<punc>>>>=|\.\.\.|===|!==|>>>|<<=|>>=|\*\*|=>|==|!=|<=|>=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|%=|\|=|&=|\^=|<<|>>|[{}()\[\];,<>+\-*/%&|^!~?:=.])
''', re.X | re.S)

ESC = {'n': '\n', 't': '\t', 'r': '\r', '0': '\0', 'b': '\b', 'f': '\f', 'v': '\v'}


def unescape(s):
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c == '\\' and i + 1 < len(s):
            n = s[i + 1]
            if n == 'u':
                if s[i + 2] == '{':
                    j = s.index('}', i)
                    out.append(chr(int(s[i + 3:j], 16)))
                    i = j + 1
                    continue
                out.append(chr(int(s[i + 2:i + 6], 16)))
                i += 6
                continue
            if n == 'x':
                out.append(chr(int(s[i + 2:i + 4], 16)))
                i += 4
                continue
            if n == '\n':
                i += 2
                continue
            out.append(ESC.get(n, n))
            i += 2
        else:
            out.append(c)
            i += 1
    # join surrogate pairs produced by 😀 style escapes
    return ''.join(out).encode('utf-16', 'surrogatepass').decode('utf-16')


def tokenize(src):
    toks, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == '`':
            # template literal: split into parts
            j, parts, cur = i + 1, [], []
            while src[j] != '`':
                if src[j] == '\\':
                    cur.append(src[j:j + 2])
                    j += 2
                elif src[j] == '$' and src[j + 1] == '{':
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == '{':
                            depth += 1
                        elif src[k] == '}':
                            depth -= 1
                        k += 1
                    parts.append(('s', unescape(''.join(cur))))
                    cur = []
                    parts.append(('e', src[j + 2:k - 1]))
                    j = k
                else:
                    cur.append(src[j])
                    j += 1
            parts.append(('s', unescape(''.join(cur))))
            toks.append(('tmpl', parts, i))
            i = j + 1
            continue
        if c == '/' and src[i + 1] not in '/*':
            # regex literal if previous significant token cannot end an expression
            prev = toks[-1] if toks else None
            if prev is None or (prev[0] == 'punc' and prev[1] not in (')', ']', '}')) or (prev[0] == 'id' and prev[1] in ('return', 'typeof')):
                j, incls = i + 1, False
                while True:
                    if src[j] == '\\':
                        j += 2
                        continue
                    if src[j] == '[':
                        incls = True
                    elif src[j] == ']':
                        incls = False
                    elif src[j] == '/' and not incls:
                        break
                    j += 1
                k = j + 1
                while k < n and src[k].isalpha():
                    k += 1
                toks.append(('regex', (src[i + 1:j], src[j + 1:k]), i))
                i = k
                continue
        m = TOKEN_RE.match(src, i)
        if not m:
            raise SyntaxError('bad token at %d: %r' % (i, src[i:i + 30]))
        i = m.end()
        kind = m.lastgroup
        if kind == 'ws':
            continue
        text = m.group(kind)
        if kind == 'num':
            v = int(text, 16) if text[:2] in ('0x', '0X') else (float(text) if any(ch in text for ch in '.eE') else int(text))
            toks.append(('num', v, m.start()))
        elif kind == 'str':
            toks.append(('str', unescape(text[1:-1]), m.start()))
        else:
            toks.append((kind, text, m.start()))
    toks.append(('eof', None, n))
    return toks


# ------------------------------------------------------------------ parser
BINPREC = {'||': 1, '&&': 2, '|': 3, '^': 4, '&': 5, '===': 6, '!==': 6, '==': 6, '!=': 6,
           '<': 7, '>': 7, '<=': 7, '>=': 7, 'instanceof': 7, 'in': 7, '<<': 8, '>>': 8, '>>>': 8,
           '+': 9, '-': 9, '*': 10, '/': 10, '%': 10, '**': 11}
ASSIGN = {'=', '+=', '-=', '*=', '/=', '%=', '|=', '&=', '^=', '<<=', '>>=', '>>>='}


class Parser:
    def __init__(self, src):
        self.src = src
        self.toks = tokenize(src)
        self.p = 0

    def peek(self, k=0):
        return self.toks[self.p + k]

    def next(self):
        t = self.toks[self.p]
        self.p += 1
        return t

    def at(self, text):
        t = self.peek()
        return t[0] in ('punc', 'id') and t[1] == text

    def eat(self, text):
        if self.at(text):
            self.p += 1
            return True
        return False

    def expect(self, text):
        if not self.eat(text):
            t = self.peek()
            line = self.src.count('\n', 0, t[2]) + 1
            raise SyntaxError('expected %r, got %r at line %d' % (text, t[1], line))

    def program(self):
        body = []
        while self.peek()[0] != 'eof':
            body.append(self.statement())
        return ('block', body)

    def block(self):
        self.expect('{')
        body = []
        while not self.at('}'):
            body.append(self.statement())
        self.expect('}')
        return ('block', body)

    def pattern(self):
        if self.eat('['):
            items = []
            while not self.at(']'):
                if self.at(','):
                    items.append(None)
                else:
                    items.append(self.pattern())
                if not self.eat(','):
                    break
            self.expect(']')
            return ('apat', items)
        if self.eat('{'):
            items = []
            while not self.at('}'):
                key = self.next()[1]
                target = ('name', key)
                if self.eat(':'):
                    target = self.pattern()
                default = None
                if self.eat('='):
                    default = self.assign()
                items.append((key, target, default))
                if not self.eat(','):
                    break
            self.expect('}')
            return ('opat', items)
        return ('name', self.next()[1])

    def statement(self):
        t = self.peek()
        if t[0] == 'punc' and t[1] == '{':
            return self.block()
        if t[0] == 'punc' and t[1] == ';':
            self.next()
            return ('empty',)
        if t[0] == 'id':
            w = t[1]
            if w in ('const', 'let', 'var'):
                self.next()
                decls = []
                while True:
                    pat = self.pattern()
                    init = self.assign() if self.eat('=') else None
                    decls.append((pat, init))
                    if not self.eat(','):
                        break
                self.eat(';')
                return ('decl', decls)
            if w == 'function' and self.peek(1)[0] == 'id':
                self.next()
                name = self.next()[1]
                params = self.params()
                body = self.block()
                return ('decl', [(('name', name), ('func', params, body, name))])
            if w == 'if':
                self.next()
                self.expect('(')
                cond = self.expr()
                self.expect(')')
                then = self.statement()
                els = self.statement() if self.eat('else') else None
                return ('if', cond, then, els)
            if w == 'for':
                self.next()
                self.expect('(')
                if self.peek()[1] in ('const', 'let', 'var') and self._is_for_of():
                    self.next()
                    pat = self.pattern()
                    kind = self.next()[1]
                    it = self.expr()
                    self.expect(')')
                    return ('forof' if kind == 'of' else 'forin', pat, it, self.statement())
                init = None if self.at(';') else self.statement_noasi()
                self.eat(';')
                cond = None if self.at(';') else self.expr()
                self.expect(';')
                upd = None if self.at(')') else self.expr()
                self.expect(')')
                return ('for', init, cond, upd, self.statement())
            if w == 'while':
                self.next()
                self.expect('(')
                cond = self.expr()
                self.expect(')')
                return ('while', cond, self.statement())
            if w == 'return':
                self.next()
                v = None
                if not self.at('}') and not self.at(';'):
                    v = self.expr()
                self.eat(';')
                return ('return', v)
            if w == 'break':
                self.next()
                self.eat(';')
                return ('break',)
            if w == 'continue':
                self.next()
                self.eat(';')
                return ('continue',)
            if w == 'throw':
                self.next()
                v = self.expr()
                self.eat(';')
                return ('throw', v)
            if w == 'try':
                self.next()
                body = self.block()
                name, handler, fin = None, None, None
                if self.eat('catch'):
                    if self.eat('('):
                        name = self.next()[1]
                        self.expect(')')
                    handler = self.block()
                if self.eat('finally'):
                    fin = self.block()
                return ('try', body, name, handler, fin)
        e = self.expr()
        self.eat(';')
        return ('expr', e)

    def _is_for_of(self):
        depth, k = 0, 1
        while True:
            t = self.peek(k)
            if t[0] == 'punc' and t[1] in '([{':
                depth += 1
            elif t[0] == 'punc' and t[1] in ')]}':
                depth -= 1
            elif depth == 0 and t[0] == 'id' and t[1] in ('of', 'in'):
                return True
            elif depth == 0 and t[0] == 'punc' and t[1] in ('=', ';'):
                return False
            k += 1

    def statement_noasi(self):
        t = self.peek()
        if t[0] == 'id' and t[1] in ('const', 'let', 'var'):
            self.next()
            decls = []
            while True:
                pat = self.pattern()
                init = self.assign() if self.eat('=') else None
                decls.append((pat, init))
                if not self.eat(','):
                    break
            return ('decl', decls)
        return ('expr', self.expr())

    def params(self):
        self.expect('(')
        ps = []
        while not self.at(')'):
            if self.eat('...'):
                ps.append(('rest', self.next()[1]))
            else:
                pat = self.pattern()
                default = self.assign() if self.eat('=') else None
                ps.append(('p', pat, default))
            if not self.eat(','):
                break
        self.expect(')')
        return ps

    def expr(self):
        e = self.assign()
        while self.at(','):
            self.next()
            e = ('seq', e, self.assign())
        return e

    def _arrow_ahead(self):
        # at '(' : find matching ')' and check for '=>'
        depth, k = 0, 0
        while True:
            t = self.peek(k)
            if t[0] == 'eof':
                return False
            if t[0] == 'punc' and t[1] in '([{':
                depth += 1
            elif t[0] == 'punc' and t[1] in ')]}':
                depth -= 1
                if depth == 0:
                    n = self.peek(k + 1)
                    return n[0] == 'punc' and n[1] == '=>'
            k += 1

    def assign(self):
        t = self.peek()
        if t[0] == 'punc' and t[1] == '(' and self._arrow_ahead():
            params = self.params()
            self.expect('=>')
            body = self.block() if self.at('{') else ('return', self.assign())
            return ('func', params, body, None)
        if t[0] == 'id' and self.peek(1)[0] == 'punc' and self.peek(1)[1] == '=>':
            name = self.next()[1]
            self.next()
            body = self.block() if self.at('{') else ('return', self.assign())
            return ('func', [('p', ('name', name), None)], body, None)
        left = self.ternary()
        t = self.peek()
        if t[0] == 'punc' and t[1] in ASSIGN:
            self.next()
            right = self.assign()
            return ('assign', t[1], left, right)
        return left

    def ternary(self):
        c = self.binary(0)
        if self.eat('?'):
            a = self.assign()
            self.expect(':')
            b = self.assign()
            return ('cond', c, a, b)
        return c

    def binary(self, minprec):
        left = self.unary()
        while True:
            t = self.peek()
            op = t[1] if t[0] in ('punc', 'id') else None
            prec = BINPREC.get(op)
            if prec is None or prec <= minprec:
                return left
            self.next()
            right = self.binary(prec if op != '**' else prec - 1)
            left = ('bin', op, left, right)

    def unary(self):
        t = self.peek()
        if t[0] == 'punc' and t[1] in ('!', '-', '+', '~'):
            self.next()
            return ('un', t[1], self.unary())
        if t[0] == 'punc' and t[1] in ('++', '--'):
            self.next()
            return ('preinc', t[1], self.unary())
        if t[0] == 'id' and t[1] in ('typeof', 'delete', 'void'):
            self.next()
            return ('un', t[1], self.unary())
        return self.postfix()

    def args(self):
        self.expect('(')
        out = []
        while not self.at(')'):
            if self.eat('...'):
                out.append(('spread', self.assign()))
            else:
                out.append(self.assign())
            if not self.eat(','):
                break
        self.expect(')')
        return out

    def postfix(self):
        t = self.peek()
        if t[0] == 'id' and t[1] == 'new':
            self.next()
            callee = self.primary()
            while self.at('.'):
                self.next()
                callee = ('member', callee, ('lit', self.next()[1]))
            args = self.args() if self.at('(') else []
            e = ('new', callee, args)
        else:
            e = self.primary()
        while True:
            t = self.peek()
            if t[0] == 'punc' and t[1] == '.':
                self.next()
                e = ('member', e, ('lit', self.next()[1]))
            elif t[0] == 'punc' and t[1] == '[':
                self.next()
                ix = self.expr()
                self.expect(']')
                e = ('member', e, ix)
            elif t[0] == 'punc' and t[1] == '(':
                e = ('call', e, self.args())
            elif t[0] == 'punc' and t[1] in ('++', '--'):
                self.next()
                e = ('postinc', t[1], e)
            elif t[0] == 'tmpl':
                raise SyntaxError('tagged templates unsupported')
            else:
                return e

    def primary(self):
        t = self.next()
        k, v = t[0], t[1]
        if k == 'num' or k == 'str':
            return ('lit', v)
        if k == 'regex':
            return ('lit', JSRegex(v[0], v[1]))
        if k == 'tmpl':
            parts = []
            for kind, s in v:
                parts.append(('lit', s) if kind == 's' else Parser(s).expr())
            return ('tmpl', parts)
        if k == 'id':
            if v == 'true':
                return ('lit', True)
            if v == 'false':
                return ('lit', False)
            if v == 'null':
                return ('lit', None)
            if v == 'undefined':
                return ('lit', undefined)
            if v == 'function':
                name = self.next()[1] if self.peek()[0] == 'id' else None
                params = self.params()
                return ('func', params, self.block(), name)
            return ('name', v)
        if k == 'punc':
            if v == '(':
                e = self.expr()
                self.expect(')')
                return e
            if v == '[':
                items = []
                while not self.at(']'):
                    if self.eat('...'):
                        items.append(('spread', self.assign()))
                    else:
                        items.append(self.assign())
                    if not self.eat(','):
                        break
                self.expect(']')
                return ('array', items)
            if v == '{':
                props = []
                while not self.at('}'):
                    if self.eat('...'):
                        props.append(('spread', self.assign()))
                    elif self.eat('['):
                        key = self.expr()
                        self.expect(']')
                        self.expect(':')
                        props.append(('kv', key, self.assign()))
                    else:
                        kt = self.next()
                        key = kt[1]
                        if kt[0] == 'num':
                            key = jsstr(key)
                        if self.at('('):   # method shorthand
                            params = self.params()
                            props.append(('kv', ('lit', key), ('func', params, self.block(), key)))
                        elif self.eat(':'):
                            props.append(('kv', ('lit', key), self.assign()))
                        else:
                            props.append(('kv', ('lit', key), ('name', key)))
                    if not self.eat(','):
                        break
                self.expect('}')
                return ('object', props)
        line = self.src.count('\n', 0, t[2]) + 1
        raise SyntaxError('unexpected token %r at line %d' % (v, line))


# ------------------------------------------------------------------ runtime helpers
def jsstr(v):
    if v is None:
        return 'null'
    if v is undefined:
        return 'undefined'
    if v is True:
        return 'true'
    if v is False:
        return 'false'
    if isinstance(v, float):
        if v.is_integer() and abs(v) < 1e21:
            return str(int(v))
        return repr(v)
    if isinstance(v, (list, tuple)):
        return ','.join('' if x is None or x is undefined else jsstr(x) for x in v)
    if isinstance(v, (bytes, bytearray)):
        return ','.join(str(b) for b in v)
    if isinstance(v, dict):
        return '[object Object]'
    if hasattr(v, 'js_unwrap'):
        return jsstr(v.js_unwrap())
    return str(v)


def truthy(v):
    v = unwrap(v)
    if v is None or v is undefined or v is False:
        return False
    if isinstance(v, (int, float)) and not isinstance(v, bool):
        return v != 0 and v == v
    if isinstance(v, str):
        return len(v) > 0
    return True


def unwrap(v):
    return v.js_unwrap() if hasattr(v, 'js_unwrap') else v


def norm(v):
    if isinstance(v, float) and v.is_integer() and abs(v) < 2 ** 53:
        return int(v)
    return v


def tonum(v):
    v = unwrap(v)
    if v is None or v is False:
        return 0
    if v is True:
        return 1
    if v is undefined:
        return float('nan')
    if isinstance(v, str):
        try:
            return norm(float(v)) if v.strip() else 0
        except ValueError:
            return float('nan')
    return v


def toint32(v):
    v = tonum(v)
    if v != v or v in (float('inf'), float('-inf')):
        return 0
    v = int(v) & 0xffffffff
    return v - (1 << 32) if v & 0x80000000 else v


def strict_eq(a, b):
    a, b = unwrap(a), unwrap(b)
    if isinstance(a, bool) != isinstance(b, bool):
        return False
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return a == b
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if a is None or a is undefined or b is None or b is undefined:
        return a is b
    return a is b


class JSFunction:
    def __init__(self, interp, params, body, env, name=None):
        self.interp, self.params, self.body, self.env, self.name = interp, params, body, env, name

    def __call__(self, *args):
        env = Env(self.env)
        for i, p in enumerate(self.params):
            if p[0] == 'rest':
                env.declare(p[1], list(args[i:]))
            else:
                v = args[i] if i < len(args) else undefined
                if v is undefined and p[2] is not None:
                    v = self.interp.ev(p[2], env)
                self.interp.bind(p[1], v, env)
        try:
            self.interp.exec(self.body, env)
        except _Return as r:
            return r.v
        return undefined


class Env:
    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def declare(self, name, v):
        self.vars[name] = v

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


class Interp:
    def __init__(self, globals_):
        self.genv = Env()
        self.genv.vars.update(globals_)
        self.protected = set()

    def run(self, src):
        self.exec(Parser(src).program(), self.genv, toplevel=True)

    # -------- statements
    def exec(self, node, env, toplevel=False):
        k = node[0]
        if k == 'block':
            benv = env if toplevel else Env(env)
            # hoist function declarations
            for st in node[1]:
                if st[0] == 'decl' and st[1][0][1] is not None and st[1][0][1][0] == 'func' and st[1][0][1][3]:
                    name = st[1][0][0][1]
                    if not (toplevel and name in self.protected):
                        benv.declare(name, self.ev(st[1][0][1], benv))
            for st in node[1]:
                self.exec(st, benv)
        elif k == 'decl':
            for pat, init in node[1]:
                if toplevel is False and pat[0] == 'name' and env is self.genv and pat[1] in self.protected:
                    continue
                if pat[0] == 'name' and pat[1] in self.protected and env.lookup(pat[1]) is self.genv and env is self.genv:
                    continue
                v = self.ev(init, env) if init is not None else undefined
                self.bind(pat, v, env)
        elif k == 'expr':
            self.ev(node[1], env)
        elif k == 'if':
            if truthy(self.ev(node[1], env)):
                self.exec(node[2], env)
            elif node[3] is not None:
                self.exec(node[3], env)
        elif k == 'for':
            fenv = Env(env)
            if node[1] is not None:
                self.exec(node[1], fenv)
            while node[2] is None or truthy(self.ev(node[2], fenv)):
                try:
                    self.exec(node[4], fenv)
                except _Break:
                    break
                except _Continue:
                    pass
                if node[3] is not None:
                    self.ev(node[3], fenv)
        elif k in ('forof', 'forin'):
            it = unwrap(self.ev(node[2], env))
            seq = list(it.keys()) if (k == 'forin' and isinstance(it, dict)) else (list(range(len(it))) if k == 'forin' else list(it))
            for v in seq:
                fenv = Env(env)
                self.bind(node[1], v, fenv)
                try:
                    self.exec(node[3], fenv)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == 'while':
            while truthy(self.ev(node[1], env)):
                try:
                    self.exec(node[2], env)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == 'return':
            raise _Return(self.ev(node[1], env) if node[1] is not None else undefined)
        elif k == 'break':
            raise _Break()
        elif k == 'continue':
            raise _Continue()
        elif k == 'throw':
            raise JSThrow(self.ev(node[1], env))
        elif k == 'try':
            try:
                self.exec(node[1], env)
            except (JSThrow, JSError) as e:
                if node[3] is None:
                    raise
                cenv = Env(env)
                if node[2]:
                    cenv.declare(node[2], e.value if isinstance(e, JSThrow) else {'name': e.kind, 'message': e.message})
                self.exec(node[3], cenv)
            finally:
                if node[4] is not None:
                    self.exec(node[4], env)
        elif k == 'empty':
            pass
        else:
            raise RuntimeError('bad statement ' + k)

    def bind(self, pat, v, env):
        if pat[0] == 'name':
            env.declare(pat[1], v)
        elif pat[0] == 'apat':
            seq = list(unwrap(v))
            for i, p in enumerate(pat[1]):
                if p is not None:
                    self.bind(p, seq[i] if i < len(seq) else undefined, env)
        else:
            for key, target, default in pat[1]:
                x = self.getmember(v, key)
                if x is undefined and default is not None:
                    x = self.ev(default, env)
                self.bind(target, x, env)

    # -------- expressions
    def ev(self, node, env):
        k = node[0]
        if k == 'lit':
            return node[1]
        if k == 'name':
            e = env.lookup(node[1])
            if e is None:
                raise JSError('ReferenceError', '%s is not defined' % node[1])
            return e.vars[node[1]]
        if k == 'tmpl':
            return ''.join(jsstr(self.ev(p, env)) for p in node[1])
        if k == 'array':
            out = []
            for it in node[1]:
                if it[0] == 'spread':
                    out.extend(list(unwrap(self.ev(it[1], env))))
                else:
                    out.append(self.ev(it, env))
            return out
        if k == 'object':
            out = {}
            for p in node[1]:
                if p[0] == 'spread':
                    src = unwrap(self.ev(p[1], env))
                    if isinstance(src, dict):
                        out.update(src)
                else:
                    key = self.ev(p[1], env)
                    out[key if isinstance(key, str) else jsstr(key)] = self.ev(p[2], env)
            return out
        if k == 'func':
            return JSFunction(self, node[1], node[2], env, node[3])
        if k == 'member':
            obj = self.ev(node[1], env)
            key = self.ev(node[2], env)
            return self.getmember(obj, key)
        if k == 'call':
            callee = node[1]
            args = []
            for a in node[2]:
                if a[0] == 'spread':
                    args.extend(list(unwrap(self.ev(a[1], env))))
                else:
                    args.append(self.ev(a, env))
            if callee[0] == 'member':
                obj = self.ev(callee[1], env)
                key = self.ev(callee[2], env)
                return self.callmethod(obj, key, args)
            f = self.ev(callee, env)
            if not callable(f):
                raise JSError('TypeError', 'not a function')
            return f(*args)
        if k == 'new':
            ctor = self.ev(node[1], env)
            args = [self.ev(a, env) for a in node[2]]
            return ctor(*args)
        if k == 'un':
            op = node[1]
            if op == 'typeof':
                try:
                    v = unwrap(self.ev(node[2], env))
                except JSError:
                    return 'undefined'
                if v is undefined:
                    return 'undefined'
                if isinstance(v, bool):
                    return 'boolean'
                if isinstance(v, (int, float)):
                    return 'number'
                if isinstance(v, str):
                    return 'string'
                if callable(v):
                    return 'function'
                return 'object'
            if op == 'delete':
                t = node[2]
                obj = unwrap(self.ev(t[1], env))
                key = self.ev(t[2], env)
                if isinstance(obj, dict):
                    obj.pop(key if isinstance(key, str) else jsstr(key), None)
                return True
            v = self.ev(node[2], env)
            if op == '!':
                return not truthy(v)
            if op == '-':
                return norm(-tonum(v))
            if op == '+':
                return norm(tonum(v))
            if op == '~':
                return ~toint32(v)
        if k == 'bin':
            op = node[1]
            if op == '&&':
                l = self.ev(node[2], env)
                return self.ev(node[3], env) if truthy(l) else l
            if op == '||':
                l = self.ev(node[2], env)
                return l if truthy(l) else self.ev(node[3], env)
            return self.binop(op, self.ev(node[2], env), self.ev(node[3], env))
        if k == 'cond':
            return self.ev(node[2], env) if truthy(self.ev(node[1], env)) else self.ev(node[3], env)
        if k == 'assign':
            op, target = node[1], node[2]
            v = self.ev(node[3], env)
            if op != '=':
                v = self.binop(op[:-1], self.ev(target, env), v)
            self.assign_to(target, v, env)
            return v
        if k in ('postinc', 'preinc'):
            old = tonum(self.ev(node[2], env))
            new = old + (1 if node[1] == '++' else -1)
            self.assign_to(node[2], new, env)
            return old if k == 'postinc' else new
        if k == 'seq':
            self.ev(node[1], env)
            return self.ev(node[2], env)
        if k == 'return':   # arrow function with expression body
            raise _Return(self.ev(node[1], env))
        raise RuntimeError('bad expression ' + k)

    def assign_to(self, target, v, env):
        if target[0] == 'name':
            e = env.lookup(target[1])
            (e or self.genv).vars[target[1]] = v
        elif target[0] == 'member':
            obj = unwrap(self.ev(target[1], env))
            key = self.ev(target[2], env)
            if isinstance(obj, list):
                if key == 'length':
                    del obj[int(v):]
                    return
                key = int(key)
                while len(obj) <= key:
                    obj.append(undefined)
                obj[key] = v
            elif isinstance(obj, dict):
                obj[key if isinstance(key, str) else jsstr(key)] = v
            elif isinstance(obj, bytearray):
                obj[int(key)] = int(v) & 0xff
            elif hasattr(obj, 'js_set'):
                obj.js_set(key, v)
            else:
                raise JSError('TypeError', 'cannot assign to member of %r' % type(obj))
        elif target[0] in ('apat', 'array'):
            raise RuntimeError('destructuring assignment unsupported')
        else:
            raise RuntimeError('bad assignment target')

    def binop(self, op, a, b):
        if op in ('===', '!=='):
            r = strict_eq(a, b)
            return r if op == '===' else not r
        a, b = unwrap(a), unwrap(b)
        if op in ('==', '!='):
            r = (a is None or a is undefined) and (b is None or b is undefined) or strict_eq(a, b)
            return r if op == '==' else not r
        if op == '+':
            if isinstance(a, str) or isinstance(b, str) or isinstance(a, (list, dict)) or isinstance(b, (list, dict)):
                return jsstr(a) + jsstr(b)
            return norm(tonum(a) + tonum(b))
        if op in ('<', '>', '<=', '>='):
            if isinstance(a, str) and isinstance(b, str):
                ka, kb = a.encode('utf-16-be', 'surrogatepass'), b.encode('utf-16-be', 'surrogatepass')
                return {'<': ka < kb, '>': ka > kb, '<=': ka <= kb, '>=': ka >= kb}[op]
            x, y = tonum(a), tonum(b)
            if x != x or y != y:
                return False
            return {'<': x < y, '>': x > y, '<=': x <= y, '>=': x >= y}[op]
        if op == '-':
            return norm(tonum(a) - tonum(b))
        if op == '*':
            return norm(tonum(a) * tonum(b))
        if op == '/':
            y = tonum(b)
            x = tonum(a)
            if y == 0:
                return float('nan') if x == 0 else math.copysign(float('inf'), x)
            return norm(x / y)
        if op == '%':
            x, y = tonum(a), tonum(b)
            return norm(math.fmod(x, y))
        if op == '**':
            return norm(tonum(a) ** tonum(b))
        if op == '&':
            return toint32(a) & toint32(b)
        if op == '|':
            return toint32(toint32(a) | toint32(b))
        if op == '^':
            return toint32(toint32(a) ^ toint32(b))
        if op == '<<':
            return toint32(toint32(a) << (toint32(b) & 31))
        if op == '>>':
            return toint32(a) >> (toint32(b) & 31)
        if op == '>>>':
            return (toint32(a) & 0xffffffff) >> (toint32(b) & 31)
        if op == 'instanceof':
            return isinstance(a, b) if isinstance(b, type) else False
        if op == 'in':
            return (a if isinstance(a, str) else jsstr(a)) in b
        raise RuntimeError('bad operator ' + op)

    def getmember(self, obj, key):
        if hasattr(obj, 'js_get'):
            return obj.js_get(key)
        if obj is None or obj is undefined:
            raise JSError('TypeError', "Cannot read property '%s' of %s" % (jsstr(key), jsstr(obj)))
        if isinstance(obj, dict):
            return obj.get(key if isinstance(key, str) else jsstr(key), undefined)
        if isinstance(obj, (list, bytes, bytearray, str, tuple)):
            if key == 'length' or key == 'byteLength':
                return len(obj)
            if isinstance(key, (int, float)) and not isinstance(key, bool):
                i = int(key)
                if 0 <= i < len(obj):
                    return obj[i]
                return undefined
            if isinstance(key, str) and key.isdigit():
                return self.getmember(obj, int(key))
            m = self.method(obj, key)
            if m is not None:
                return m
            return undefined
        if isinstance(obj, JSFunction):
            return undefined
        v = getattr(obj, key, undefined) if isinstance(key, str) else undefined
        return v

    def callmethod(self, obj, key, args):
        if hasattr(obj, 'js_call'):
            return obj.js_call(key, args)
        f = self.getmember(obj, key)
        if not callable(f):
            raise JSError('TypeError', '%s is not a function' % jsstr(key))
        return f(*args)

    def method(self, obj, name):
        I = self
        if isinstance(obj, list):
            def sort(cmp=None):
                import functools
                if cmp is None or cmp is undefined:
                    obj.sort(key=lambda x: jsstr(x).encode('utf-16-be', 'surrogatepass'))
                else:
                    obj.sort(key=functools.cmp_to_key(lambda a, b: (lambda r: -1 if r < 0 else (1 if r > 0 else 0))(tonum(cmp(a, b)))))
                return obj

            def splice(start, count=None, *items):
                start = int(start)
                count = len(obj) - start if count is None else int(count)
                removed = obj[start:start + count]
                obj[start:start + count] = list(items)
                return removed

            def fill(v):
                for i in range(len(obj)):
                    obj[i] = v
                return obj
            table = {
                'push': lambda *a: (obj.extend(a), len(obj))[1],
                'pop': lambda: obj.pop() if obj else undefined,
                'shift': lambda: obj.pop(0) if obj else undefined,
                'unshift': lambda *a: (obj.__setitem__(slice(0, 0), list(a)), len(obj))[1],
                'map': lambda f: [f(x, i) if _arity(f) > 1 else f(x) for i, x in enumerate(list(obj))],
                'filter': lambda f: [x for x in obj if truthy(f(x))],
                'forEach': lambda f: ([f(x) for x in list(obj)], undefined)[1],
                'find': lambda f: next((x for x in obj if truthy(f(x))), undefined),
                'findIndex': lambda f: next((i for i, x in enumerate(obj) if truthy(f(x))), -1),
                'some': lambda f: any(truthy(f(x)) for x in obj),
                'every': lambda f: all(truthy(f(x)) for x in obj),
                'includes': lambda v: any(strict_eq(x, v) for x in obj),
                'indexOf': lambda v: next((i for i, x in enumerate(obj) if strict_eq(x, v)), -1),
                'slice': lambda a=0, b=None: obj[int(a):(None if b is None else int(b))],
                'concat': lambda *a: obj + [y for x in a for y in (x if isinstance(x, list) else [x])],
                'join': lambda sep=',': sep.join(jsstr(x) for x in obj),
                'reverse': lambda: (obj.reverse(), obj)[1],
                'sort': sort, 'splice': splice, 'fill': fill,
                'reduce': lambda f, init=undefined: _reduce(f, obj, init),
                'keys': lambda: list(range(len(obj))),
            }
            return table.get(name)
        if isinstance(obj, str):
            table = {
                'slice': lambda a=0, b=None: obj[int(a):(None if b is None else int(b))],
                'substring': lambda a=0, b=None: obj[int(a):(None if b is None else int(b))],
                'split': lambda sep: list(obj) if sep == '' else obj.split(sep),
                'charCodeAt': lambda i=0: ord(obj[int(i)]),
                'indexOf': lambda s: obj.find(s),
                'includes': lambda s: s in obj,
                'startsWith': lambda s: obj.startswith(s),
                'repeat': lambda n: obj * int(n),
                'toString': lambda: obj,
                'match': lambda r: ([obj] if r.test(obj) else None),
                'padStart': lambda n, c=' ': obj.rjust(int(n), c),
                'toUpperCase': lambda: obj.upper(), 'toLowerCase': lambda: obj.lower(),
            }
            return table.get(name)
        if isinstance(obj, (bytes, bytearray)):
            table = {
                'subarray': lambda a=0, b=None: obj[int(a):(None if b is None else int(b))],
                'slice': lambda a=0, b=None: obj[int(a):(None if b is None else int(b))],
                'fill': lambda v: bytearray([int(v)] * len(obj)),
            }
            return table.get(name)
        return None


def _arity(f):
    if isinstance(f, JSFunction):
        return len(f.params)
    return 1


def _reduce(f, seq, init):
    acc, start = init, 0
    if init is undefined:
        acc, start = seq[0], 1
    for x in seq[start:]:
        acc = f(acc, x)
    return acc
